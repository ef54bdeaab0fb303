"""Video-level full-reference scoring: shard frame pairs over the GPUs of one box, score them with the
B200 kernels, gather the per-frame scores with ONE collective and pool per video.

Replaces the host loops of the reference's offline scorers (prep.py:171-216, test2_prep.py:132-193: per
video -> per batch -> model(ref, render) -> np.mean/std/min/max) and the pooling / correlation part of
``MetricCollectionLogger`` (nerf_qa/logger.py:93-141).  Frames are independent, so the data path has no
collective; the only exchange is the score gather.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(total: int, rank: int, world: int) -> np.ndarray:
    """Round-robin frame assignment ``g % world == rank``: balances videos of unequal length."""
    return np.arange(rank, total, world, dtype=np.int64)


def padded_shard_len(total: int, world: int) -> int:
    return (total + world - 1) // world


def video_offsets(lengths: Sequence[int]) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.asarray(lengths, dtype=np.int64))
    return off


def frame_to_video(lengths: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """global frame index -> (video id, frame id inside the video)."""
    off = video_offsets(lengths)
    vid = np.repeat(np.arange(len(lengths), dtype=np.int64), np.asarray(lengths, dtype=np.int64))
    frame = np.arange(off[-1], dtype=np.int64) - off[vid]
    return vid, frame


def gather_scores(local: torch.Tensor, total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather the round-robin shards (one collective, equal counts, NaN padding) and undo the
    interleave.  ``local`` lives on the device NCCL uses (or on the CPU for gloo)."""
    if world == 1:
        return local[:total].clone()
    n = padded_shard_len(total, world)
    send = torch.full((n,), float("nan"), dtype=torch.float32, device=local.device)
    send[: local.numel()] = local
    recv = torch.empty(world * n, dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    # recv[r*n + i] is global frame i*world + r
    return recv.view(world, n).t().reshape(-1)[:total].contiguous()


def pool_videos(scores: np.ndarray, lengths: Sequence[int]) -> Dict[str, np.ndarray]:
    """Per-video mean / std / min / max of the frame scores (prep.py:191-198; population std like np.std)."""
    off = video_offsets(lengths)
    s = np.asarray(scores, dtype=np.float64)
    out = {k: np.empty(len(lengths)) for k in ("mean", "std", "min", "max")}
    for v in range(len(lengths)):
        seg = s[off[v]: off[v + 1]]
        out["mean"][v], out["std"][v], out["min"][v], out["max"][v] = seg.mean(), seg.std(), seg.min(), seg.max()
    return out


def correlations(pred: np.ndarray, target: np.ndarray) -> Dict[str, float]:
    """PLCC / SRCC / KTCC over videos (logger.py:93-102)."""
    import scipy.stats as st
    return {"plcc": float(st.pearsonr(pred, target)[0]), "srcc": float(st.spearmanr(pred, target)[0]),
            "ktcc": float(st.kendalltau(pred, target)[0])}


REAL_SCENES = ("train", "m60", "playground", "truck", "fortress", "horns", "trex", "room")          # logger.py:217
SYNTH_SCENES = ("ship", "lego", "drums", "ficus", "hotdog", "materials", "mic", "chair")            # logger.py:218


def summarize(frame_scores: np.ndarray, video_ids: Sequence, scene_ids: Sequence, mos: Optional[np.ndarray] = None,
              name: str = "score") -> Dict[str, object]:
    """The wandb-free content of ``MetricCollectionLogger.log_summary`` (logger.py:144-282) for one metric:

    * per-video mean of the frame scores (logger.py:157-160) and per-scene mean of those video means (logger.py:173-178);
    * their averages over videos / over scenes (logger.py:182-186);
    * with ``mos`` (one value per FRAME, constant inside a video, exactly how the reference logs it): PLCC / SRCC / KTCC
      over all videos (logger.py:264-269), per scene with at least two videos (logger.py:230-231), over the videos of
      the real and of the synthetic scenes (logger.py:239-249), and min / mean over scenes of the absolute per-scene
      correlations (logger.py:252-266).

    ``video_ids`` / ``scene_ids`` have one entry per frame; videos and scenes are reported in ``np.unique`` order."""
    s = np.asarray(frame_scores, dtype=np.float64)
    vids, sids = np.asarray(video_ids), np.asarray(scene_ids)
    uv, us = np.unique(vids), np.unique(sids)
    vmean, vscene, vmos = {}, {}, {}
    for v in uv:
        m = vids == v
        vmean[v] = float(np.mean(s[m]))
        vscene[v] = sids[m][0]
        if mos is not None:
            vmos[v] = float(np.mean(np.asarray(mos, dtype=np.float64)[m]))
    scene_videos = {sc: [v for v in uv if vscene[v] == sc] for sc in us}
    smean = {sc: float(np.mean([vmean[v] for v in scene_videos[sc]])) for sc in us}
    out: Dict[str, object] = {
        "name": name, "videos": list(uv), "scenes": list(us),
        "video_mean": np.array([vmean[v] for v in uv]), "scene_mean": np.array([smean[sc] for sc in us]),
        "average_over_videos": float(np.mean([vmean[v] for v in uv])),
        "average_over_scenes": float(np.mean([smean[sc] for sc in us])),
    }
    if mos is None:
        return out
    pred = np.array([vmean[v] for v in uv])
    gt = np.array([vmos[v] for v in uv])
    out["video_mos"] = gt
    if len(uv) > 1:
        out["correlations"] = correlations(pred, gt)
    per_scene = {}
    groups = {"real": ([], []), "synthetic": ([], [])}
    for sc in us:
        p = np.array([vmean[v] for v in scene_videos[sc]])
        g = np.array([vmos[v] for v in scene_videos[sc]])
        if len(p) > 1:
            per_scene[sc] = correlations(p, g)
        grp = "real" if str(sc) in REAL_SCENES else "synthetic" if str(sc) in SYNTH_SCENES else None
        if grp:
            groups[grp][0].append(p)
            groups[grp][1].append(g)
    out["scene_correlations"] = per_scene
    for grp, (ps, gs) in groups.items():
        if len(ps) > 1:
            out[f"{grp}_correlations"] = correlations(np.concatenate(ps), np.concatenate(gs))
    if per_scene:
        for k in ("plcc", "srcc", "ktcc"):
            vals = np.abs([c[k] for c in per_scene.values()])
            out[f"scene_min_{k}"] = float(np.min(vals))
            out[f"scene_mean_{k}"] = float(np.mean(vals))
    return out


@dataclass
class VideoScores:
    per_frame: np.ndarray            # [total] in global frame order
    per_video: Dict[str, np.ndarray]  # mean/std/min/max, each [num_videos]


FrameSource = Callable[[np.ndarray], Tuple[torch.Tensor, torch.Tensor]]
ScoreFn = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]


def score_videos(score_fn: ScoreFn, frames: FrameSource, lengths: Sequence[int], batch_size: int = 32,
                 rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 device: Optional[torch.device] = None) -> VideoScores:
    """Score ``sum(lengths)`` frame pairs.

    ``frames(global_indices) -> (ref, dist)`` returns the [b,3,H,W] batches for the given global frame ids
    (decode / render / synthesise them on the fly); ``score_fn(ref, dist) -> [b]`` is e.g. a
    ``nerf_qa_b200.DISTS`` module.  With an initialised process group each rank scores its round-robin shard
    and one all-gather assembles the result on every rank."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if world > 1 else 0
    total = int(np.sum(lengths))
    mine = shard_indices(total, rank, world)
    chunks: List[torch.Tensor] = []
    for i in range(0, len(mine), batch_size):
        ref, dst = frames(mine[i: i + batch_size])
        with torch.no_grad():
            chunks.append(score_fn(ref, dst).reshape(-1).to(torch.float32))
    if chunks:
        local = torch.cat(chunks)
    else:
        local = torch.empty(0, dtype=torch.float32, device=device if device is not None else "cpu")
    if device is not None:
        local = local.to(device)
    full = gather_scores(local, total, rank, world, group)
    per_frame = full.detach().cpu().numpy()
    return VideoScores(per_frame=per_frame, per_video=pool_videos(per_frame, lengths))
