"""``torch.ops.nerfqa_b200.*``: the thin PyTorch C++ extension over the C-ABI (csrc_torch/nqa_torch_ops.cpp).

``load()`` loads ``_nqa_torch_ops.so`` (built in-tree by ``make``) and registers fake (meta) kernels, so that
``torch.compile`` / ``torch.export`` can trace through the scoring ops.  The extension holds no arithmetic: each operator
is one C-ABI call on torch's current stream.  ``_lib.Context`` routes the no-grad scoring calls through these operators
when the extension is present (``NQA_BINDING=ctypes`` keeps the raw ctypes binding, the two are bit-identical)."""
from __future__ import annotations

import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(_HERE, "_nqa_torch_ops.so")
NUM_CHANNELS = 1475
_CHNS = (64, 128, 256, 512, 512)
OPS = ("ctx_create", "ctx_destroy", "dists_forward", "adists_forward", "dists_adists_forward", "dists_features", "l2pool")

_ops = None


def available() -> bool:
    return os.path.exists(EXT_PATH)


def _register_fakes() -> None:
    def fake(name):
        return torch.library.register_fake(f"nerfqa_b200::{name}")

    @fake("dists_forward")
    def _(ctx, x, y, precision, want_s1s2):
        b = x.shape[0]
        return x.new_empty((b,), dtype=torch.float32), x.new_empty((b if want_s1s2 else 0, 2, NUM_CHANNELS), dtype=torch.float32)

    @fake("adists_forward")
    def _(ctx, x, y, window, precision):
        return x.new_empty((x.shape[0],), dtype=torch.float32)

    @fake("dists_adists_forward")
    def _(ctx, x, y, window, precision):
        return x.new_empty((x.shape[0],), dtype=torch.float32), x.new_empty((x.shape[0],), dtype=torch.float32)

    @fake("dists_features")
    def _(ctx, x, precision):
        b, _, h, w = x.shape
        outs = []
        for k, c in enumerate(_CHNS):
            if k > 0:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            outs.append(x.new_empty((b, c, h, w), dtype=torch.float32))
        return outs

    @fake("l2pool")
    def _(ctx, x):
        b, c, h, w = x.shape
        return x.new_empty((b, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), dtype=torch.float32)


def load():
    """The ``torch.ops.nerfqa_b200`` namespace (loads the extension once).  Raises when it has not been built."""
    global _ops
    if _ops is not None:
        return _ops
    if not available():
        raise RuntimeError(f"{EXT_PATH} not found: build it with `make`")
    from . import _lib
    _lib.load_library()                      # the C-ABI library first: the extension links against it
    torch.ops.load_library(EXT_PATH)
    _register_fakes()
    _ops = torch.ops.nerfqa_b200
    return _ops


def load_if_enabled() -> Optional[object]:
    """What ``_lib.Context`` uses: the operators unless ``NQA_BINDING=ctypes`` or the extension is not built."""
    if os.environ.get("NQA_BINDING", "torch") == "ctypes" or not available():
        return None
    return load()
