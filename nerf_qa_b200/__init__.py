"""nerf_qa_b200 — B200-native DISTS / A-DISTS frame-pair scorers (drop-in for the full-reference hot
path of kobejean/nerf-qa).  All scoring runs through the C-ABI library ``libnerfqa_b200.so`` (hand-written
sm_100a kernels); there is no CPU or eager-PyTorch fallback."""
from ._lib import NerfQAError, load_library  # noqa: F401
from .dists import DISTS, DISTSOriginal, DISTSSoftmax, L2pooling, prepare_image  # noqa: F401
from .adists import ADISTS, Downsample  # noqa: F401
from .fr_model import NeRFQAModel  # noqa: F401
from . import synth, video  # noqa: F401

__all__ = ["DISTS", "DISTSOriginal", "DISTSSoftmax", "ADISTS", "NeRFQAModel", "L2pooling", "Downsample", "prepare_image", "NerfQAError", "load_library", "synth", "video"]
