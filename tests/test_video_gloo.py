"""Host logic of the multi-GPU path on CPU: round-robin sharding, the single score all-gather (gloo,
world_size 2) and per-video pooling.  The scorer is the CPU oracle here (tests may use it)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_qa_b200 import video

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_score(ref, dst):
    # cheap deterministic stand-in with the scorer's signature: [b,3,H,W] x2 -> [b]
    return (ref - dst).abs().mean(dim=(1, 2, 3))


def _frames(idx):
    g = torch.Generator().manual_seed(1)
    base = torch.rand(23, 3, 8, 8, generator=g)
    idx = torch.as_tensor(np.asarray(idx))
    return base[idx], base[(idx * 7 + 3) % 23]


LENGTHS = [6, 9, 8]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = video.score_videos(_fake_score, _frames, LENGTHS, batch_size=4)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), res.per_frame)
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_layout():
    total = 11
    for world in (1, 2, 3, 4, 8):
        seen = np.concatenate([video.shard_indices(total, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(total))
        assert video.padded_shard_len(total, world) * world >= total
    vid, frame = video.frame_to_video([2, 3])
    assert vid.tolist() == [0, 0, 1, 1, 1] and frame.tolist() == [0, 1, 0, 1, 2]


def test_pooling_matches_numpy():
    s = np.random.RandomState(0).rand(sum(LENGTHS))
    p = video.pool_videos(s, LENGTHS)
    off = video.video_offsets(LENGTHS)
    for v in range(3):
        seg = s[off[v]:off[v + 1]]
        assert p["mean"][v] == pytest.approx(seg.mean()) and p["std"][v] == pytest.approx(seg.std())
        assert p["min"][v] == seg.min() and p["max"][v] == seg.max()


def test_two_rank_gather_equals_single_process(tmp_path):
    single = video.score_videos(_fake_score, _frames, LENGTHS, batch_size=5, rank=0, world=1)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(r0, r1)                     # every rank holds the gathered scores
    np.testing.assert_array_equal(r0, single.per_frame)       # bit-identical to the single-process run
    assert not np.isnan(r0).any()


def test_video_correlations_follow_pooled_means():
    r = np.random.RandomState(3)
    lengths = [5, 7, 6, 9, 4, 8]
    per_frame = np.concatenate([m + 0.01 * r.randn(n) for m, n in zip([0.1, 0.3, 0.2, 0.5, 0.4, 0.6], lengths)])
    pooled = video.pool_videos(per_frame, lengths)
    mos = np.array([4.5, 3.0, 3.9, 1.5, 2.2, 1.0])
    c = video.correlations(pooled["mean"], mos)
    assert c["srcc"] == pytest.approx(-1.0) and c["plcc"] < -0.9 and c["ktcc"] == pytest.approx(-1.0)
