#!/usr/bin/env python
"""Benchmark of the DISTS / A-DISTS frame-pair hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dists|adists] [--precision bf16|bf16x3|fp32]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of synthetic frame pairs:
  dists  (default, BASELINE configs[1]): DISTS on 64 synthetic 1920x1080 frame pairs per GPU;
  adists (configs[2]):                   A-DISTS on 32 synthetic 1920x1080 frame pairs per GPU.
With N > 1 (torchrun) every rank scores its own batch (weak scaling, frames are independent) and the
per-frame scores are all-gathered once per step -- the only collective on this path.

One JSON line is printed by rank 0.  `value` is whole-job frame-pairs/s with inputs resident in HBM
(CUDA events, max over ranks); `e2e` is the same metric through the host-buffer C-ABI call with the
H2D copies of the frames and the D2H read of the scores inside the timed region; `roofline` is the
dominant kernel family (the tcgen05 convolutions) from CUDA events recorded around every launch during
the timed region; `cpu_baseline` is the oracle port of the reference timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W = 1080, 1920
WORKLOADS = {
    "dists": dict(pairs=64, metric="dists_1080p_frame_pairs_per_sec",
                  desc="DISTS batch of 64 synthetic 1920x1080 frame pairs per GPU (BASELINE configs[1])"),
    "adists": dict(pairs=32, metric="adists_1080p_frame_pairs_per_sec",
                   desc="A-DISTS batch of 32 synthetic 1920x1080 frame pairs per GPU (BASELINE configs[2])"),
    "train": dict(pairs=32, metric="dists_512_fwd_bwd_frame_pairs_per_sec", hw=(512, 512),
                  desc="DISTS forward+backward (require_grad=True) on 32 synthetic 512x512 crop pairs per GPU "
                       "(BASELINE configs[4]); a step = scores + d(mean score)/d frames"),
}
# algorithmic work per 1080p frame PAIR
TRUNK_FLOP_PER_PAIR = 2.5386e12


def load_conv_traffic():
    """DRAM bytes per launch of the trunk-convolution family from the committed `ncu --set full` capture of the same
    bench command (profiles/r1z_ncu_full_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of every
    stage1 / pair / halo kernel of one 16-frame micro-batch, averaged over those launches)."""
    path = os.path.join(ROOT, "profiles", "r1z_ncu_full_summary.json")
    try:
        with open(path) as f:
            d = json.load(f)
        cols = d["columns"]
        kn, rd, wr = cols.index("Kernel Name"), cols.index("dram__bytes_read.sum"), cols.index("dram__bytes_write.sum")
        rows = [r for r in d["rows"] if any(t in r[kn] for t in ("stage1_fused", "conv3x3_tc_pair", "conv3x3_tc_halo"))]
        # the capture covers a little more than one micro-batch: keep the first 12 conv launches (1 fused + 11)
        rows = rows[:12]
        if len(rows) < 12:
            return None
        total_gb = sum(float(r[rd]) + float(r[wr]) for r in rows)
        return {"bytes_per_launch": total_gb * 1e9 / len(rows), "launches_in_capture": len(rows),
                "source": "profiles/r1z_ncu_full_summary.json (ncu --set full, bench.py --pairs 8, 16 frames per launch)"}
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_frames_device(pairs: int, device, seed: int = 0, H: int = H, W: int = W):
    """Synthetic "NeRF-like" frames generated on the device: box-blurred noise + the five distortion
    types of nerf_qa_b200.synth, 8 distinct pairs tiled to the batch."""
    import torch.nn.functional as F
    g = torch.Generator(device=device).manual_seed(seed)
    distinct = min(8, pairs)

    def smooth(n):
        r = torch.rand(n, 3, H, W, generator=g, device=device)
        r = F.avg_pool2d(F.pad(r, (4, 4, 4, 4), mode="reflect"), 9, stride=1)
        lo, hi = r.amin(dim=(1, 2, 3), keepdim=True), r.amax(dim=(1, 2, 3), keepdim=True)
        return (r - lo) / (hi - lo)

    ref = smooth(distinct)
    dist = torch.empty_like(ref)
    for i in range(distinct):
        k = i % 5
        if k == 0:
            d = ref[i] + 0.05 * torch.randn(ref[i].shape, generator=g, device=device)
        elif k == 1:
            d = ref[i] + 0.2 * torch.randn(ref[i].shape, generator=g, device=device)
        elif k == 2:
            d = F.avg_pool2d(F.pad(ref[i:i + 1], (3, 3, 3, 3), mode="reflect"), 7, stride=1)[0]
        elif k == 3:
            d = torch.roll(ref[i], 5, dims=2)
        else:
            d = smooth(1)[0]
        dist[i] = d.clamp(0, 1)
    reps = (pairs + distinct - 1) // distinct
    return ref.repeat(reps, 1, 1, 1)[:pairs].contiguous(), dist.repeat(reps, 1, 1, 1)[:pairs].contiguous()


def build_model(kind: str, precision: str, device):
    import nerf_qa_b200
    from nerf_qa_b200 import synth
    ws, bs = synth.make_vgg_weights(1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.DISTS(precision=precision) if kind == "dists" else nerf_qa_b200.ADISTS(precision=precision)
    sd = m.state_dict()
    for i, idx in enumerate(synth.VGG_FEATURE_IDX):
        sd[f"stage{synth.CONV_STAGE[i]}.{idx}.weight"] = ws[i]
        sd[f"stage{synth.CONV_STAGE[i]}.{idx}.bias"] = bs[i]
    m.load_state_dict(sd)
    return m.to(device).eval()


def cpu_reference_pairs_per_sec(kind: str, steps: int, warmup: int):
    """The reference's CPU path (oracle port: same ATen ops, all host threads) on a bounded sample:
    one 1080p frame pair per step."""
    from nerf_qa_b200 import synth
    from oracle import reference_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ws, bs = synth.make_vgg_weights(1234)
    ab = np.load(os.path.join(ROOT, "nerf_qa_b200", "data", "dists_alpha_beta.npy"))
    a, b = torch.from_numpy(ab[0].copy()), torch.from_numpy(ab[1].copy())
    x, y = synth.make_frame_pairs(1, H, W, seed=0)

    def step():
        with torch.no_grad():
            return O.dists_score(x, y, ws, bs, a, b) if kind == "dists" else O.adists_score(x, y, ws, bs)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    # a 1080p pair costs ~5-30 s of CPU: bound the run to a few minutes whatever K/W the driver passes
    steps = max(1, min(args.steps, 6 if args.workload == "dists" else 2))
    warmup = max(0, min(args.warmup, 1))
    pps, dt, cores = cpu_reference_pairs_per_sec(args.workload, steps, warmup)
    sample = f"1 synthetic 1920x1080 frame pair per step, {steps} timed steps, {warmup} warm-up, torch CPU fp32"
    line = {
        "impl": "reference", "metric": wl["metric"], "value": pps, "unit": "frame-pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "resolution": "1920x1080", "pairs_per_step": 1,
                   "note": "reference CPU path = oracle/reference_oracle.py (torch-CPU port pinned to the reference "
                           "modules by tests/golden); the reference itself is pure Python and cannot be pip-installed "
                           "offline (needs torchvision ImageNet weights download)"},
        "cpu_baseline": {"value": pps, "unit": "frame-pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dists", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--pairs", type=int, default=0, help="override frame pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    from nerf_qa_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (nerf_qa_b200 has no CPU fallback)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    wl = WORKLOADS[args.workload]
    pairs = args.pairs or wl["pairs"]
    steps, warmup = args.steps, max(args.warmup, 3)
    peaks = load_peaks()
    H, W = wl.get("hw", (1080, 1920))
    train = args.workload == "train"

    model = build_model("adists" if args.workload == "adists" else "dists", args.precision, device)
    x, y = make_frames_device(pairs, device, seed=rank, H=H, W=W)
    if train:
        x.requires_grad_(True)
        y.requires_grad_(True)
    ctx = model._context(device)
    gathered = torch.empty(world * pairs, device=device, dtype=torch.float32) if world > 1 else None

    def step():
        if train:
            x.grad = y.grad = None
            s = model(x, y, require_grad=True)
            s.mean().backward()
            s = s.detach()
        else:
            with torch.no_grad():
                s = model(x, y) if args.workload == "dists" else model(x, y, as_loss=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)     # the path's one collective: per-frame score gather
            return gathered
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        scores = step()
    barrier()
    if not bool(torch.isfinite(scores).all()):
        raise SystemExit("bench.py: non-finite scores")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count()
    ctx.profile_read()
    ctx.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        scores = step()
    ev1.record()
    barrier()
    ctx.profile_enable(False)
    prof = ctx.profile_read()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    lsum = torch.tensor([float(launches)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    ms_step = float(ms_total.item()) / steps
    value = world * pairs / (ms_step / 1e3)

    # ---- end to end through the host-buffer C-ABI call (H2D of the frames + D2H of the scores inside) ----
    e2e = None
    if args.workload == "dists" and not args.no_e2e:
        xh, yh = x.detach().cpu().pin_memory(), y.detach().cpu().pin_memory()
        out = torch.empty(pairs, dtype=torch.float32).pin_memory()
        model.score_host(xh, yh, device)                     # warm-up (staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.dists_forward_host(xh, yh, model._prec(), out=out)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * pairs * steps / float(dt.item()), "unit": "frame-pairs/s",
               "h2d_bytes_per_step": int(2 * pairs * 3 * H * W * 4), "d2h_bytes_per_step": int(pairs * 4),
               "api": "DISTS.score_host -> nqa_dists_forward_host (pinned host frames, copy stream overlapped with compute)",
               "matches_device_path": bool(torch.equal(out, scores[rank * pairs:(rank + 1) * pairs].cpu()
                                                       if world > 1 else scores.cpu()))}

    if rank == 0:
        # the trunk's 13 convolutions: fused stage 1 (normalise + conv1_1 + conv1_2) + the 11 other tcgen05 convs
        # (+ the stand-alone conv1_1 kernel when the fused stage-1 kernel is switched off)
        conv_fams = [prof[k] for k in ("conv_tc", "stage1", "conv1_1") if k in prof]
        tc_ms, tc_launches = sum(v[0] for v in conv_fams), sum(v[1] for v in conv_fams)
        # train: forward, recomputed forward inside the backward call, and 12 of 13 data-gradient convolutions
        flop_pair = TRUNK_FLOP_PER_PAIR * (H * W) / (1080 * 1920) * (3.0 if train else 1.0)
        tc_flop = flop_pair * pairs * steps
        tc_peak = peaks["bf16_tflops_sustained"] * (1.0 if args.precision == "bf16" else 1.0)
        achieved = tc_flop / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else None
        mma_factor = 3.0 if args.precision == "bf16x3" else 1.0
        total_prof_ms = sum(v[0] for v in prof.values())
        roofline = {
            "kernel": "VGG16 trunk convolutions on tcgen05: stage1_fused_kernel (normalise + conv1_1 + conv1_2) + "
                      "conv3x3_tc[_halo]_kernel (11 layers), bias + ReLU fused",
            "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
            "frac": (achieved / tc_peak) if achieved else None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernel timed inside a long step",
            "algorithmic_flop_per_launch": tc_flop / max(tc_launches, 1),
            "issued_mma_flop_factor": mma_factor,
            "avg_launch_ms": tc_ms / max(tc_launches, 1), "launches": tc_launches,
            "share_of_step": tc_ms / total_prof_ms if total_prof_ms > 0 else None,
            "traffic": None,
        }
        tr = load_conv_traffic() if (args.workload == "dists" and args.precision == "bf16") else None
        if tr is not None:
            roofline["traffic"] = tr["bytes_per_launch"]
            roofline["traffic_source"] = tr["source"]
            # algorithmic HBM bytes of the same 12 launches: frames in (fp32), every conv output written once (bf16),
            # every conv input except the fused stage's read once
            roofline["traffic_note"] = "DRAM bytes per conv launch (avg of the 12 trunk launches of one 8-pair micro-batch)"
        families = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in prof.items() if v[1]}
        line = {
            "metric": wl["metric"], "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["desc"], "pairs_per_gpu": pairs, "resolution": f"{W}x{H}",
                       "trunk_precision": args.precision, "vgg16": "random-init (seed 1234), shipped alpha/beta",
                       "micro_batch_pairs": ctx.micro_batch(pairs, H, W, model._prec()),
                       "l2": f"inputs are {2 * pairs * 3 * H * W * 4 / 1e9:.2f} GB per step, larger than the 126 MB L2",
                       "parallelism": f"frame-sharded dp{world}, one all_gather of scores per step" if world > 1 else "single GPU"},
            "e2e": e2e, "gpu_launches": int(lsum.item()), "clocks": clocks, "roofline": roofline,
            "kernel_families": families,
        }
        if world == 1 and not args.no_cpu_baseline and not train:
            n = 2 if args.workload == "dists" else 1
            pps, dt_cpu, cores = cpu_reference_pairs_per_sec(args.workload, n, 1 if args.workload == "dists" else 0)
            line["cpu_baseline"] = {"value": pps, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} timed 1920x1080 frame pair(s), one per step, oracle port of the "
                                              f"reference (torch CPU fp32, {cores} threads)"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        # Every rank is done and rank 0 has printed its line.  Leave without the NCCL communicator teardown:
        # destroy_process_group() was observed to block for minutes under torchrun on the B200 box after the line
        # had been printed, and the driver times the whole process.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
