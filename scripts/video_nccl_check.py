"""BASELINE configs[3] at reduced size on real GPUs: synthetic scenes x frames, frame pairs sharded round-robin over the
ranks (torchrun, NCCL), one all-gather of the per-frame scores, temporal mean pooling.  Rank 0 also scores every frame
alone and checks that the gathered result is bit-identical (the reduction order of a frame never depends on the batch
or rank it travelled in).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/video_nccl_check.py
"""
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_qa_b200 import synth, video  # noqa: E402
from tests._util import make_dists  # noqa: E402

SCENES, FRAMES, H, W = 4, 24, 270, 480


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warnings.simplefilter("ignore")
    model = make_dists("bf16", device=str(dev))
    lengths = [FRAMES] * SCENES

    def frames(idx):
        # frame g of the job: generated from its global id, whichever rank asks for it
        xs, ys = zip(*[synth.make_frame_pairs(1, H, W, seed=1000 + int(g)) for g in idx])
        return torch.cat(xs).to(dev), torch.cat(ys).to(dev)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = video.score_videos(model, frames, lengths, batch_size=8, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        alone = video.score_videos(model, frames, lengths, batch_size=5, rank=0, world=1, device=dev)
        same = bool(np.array_equal(alone.per_frame, res.per_frame))
        print(f"world {world}: {SCENES}x{FRAMES} frames {W}x{H} in {dt:.2f} s; per-video mean {np.round(res.per_video['mean'], 5)}; "
              f"gathered == single-rank scores: {same}", flush=True)
        assert same and res.per_frame.shape == (SCENES * FRAMES,) and not np.isnan(res.per_frame).any()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
