#!/usr/bin/env python
"""bench.py — image-pairs/sec of the pairwise deep front-end hot path (detect + match + verify).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

Workload (config.workload): the steady state of BASELINE.json configs[3] with the deep_front_end.yaml matcher —
a synthetic 640x480 frame sequence, `Sequential(max_frame_lookahead=20)` pairs, SuperPoint (max 5000 keypoints) ->
LightGlue -> RANSAC-5pt essential matrix.  One STEP = 2 new frames arriving: 2 detections + 40 pair matches + 40
verifications against the 20-frame window whose features are resident.  Pairs shard across GPUs with no data-path
collective (one weight broadcast at start-up), per-GPU work is fixed => "weak" scaling.

`value` times the device-resident path (frames already in HBM, features/matches stay in HBM, only per-pair scalars
come back); `e2e` times the same step through the GTSfM plugin classes with HOST numpy buffers, so every H2D / D2H copy
the per-call API implies is inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import deque
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W = 480, 640
MAX_KP = 5000
LOOKAHEAD = 20
NEW_FRAMES = 2
PAIRS_PER_STEP = LOOKAHEAD * NEW_FRAMES
THR_PX = 4.0
MATCH_BATCH = 8  # pairs per lock-step LightGlue batch (the library maximum)
DOMINANT_KERNEL = "k_flash"  # prefix: k_flash_ps / k_flash_ts / k_flash_ws / k_flash_tc (tcgen05) or k_flash_attn (forced SIMT)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_flash_ps launch at this workload, from the `ncu --set full` capture
# summarised in profiles/r01_flash_ps.txt (30.80 MB read + 0.68 MB written; the algorithmic minimum - q, k, v, o planes
# of both images once - is 4 x 2 x 5000 x 256 x 4 B = 41 MB, i.e. K / V re-reads are served by L2)
DOMINANT_KERNEL_DRAM_BYTES_PER_LAUNCH = 30.80e6 + 0.68e6
CONFIG = {
    "workload": "SuperPoint+LightGlue+RANSAC-5pt, synthetic 640x480 sequence, Sequential lookahead 20 (BASELINE configs[3] steady state, deep_front_end.yaml matcher)",
    "frame": [H, W], "max_keypoints": MAX_KP, "lookahead": LOOKAHEAD, "new_frames_per_step": NEW_FRAMES,
    "pairs_per_step": PAIRS_PER_STEP, "lightglue": "9 layers, full depth (synthetic 'bench' weights: no early exit, nothing pruned)",
    "ransac": "5pt, 1000 hypotheses, thr 4 px, conf 0.999999", "weights": "seeded synthetic (no checkpoint offline)",
    "l2": "256 MiB flush between timed steps", "parallelism": "pairs sharded per GPU, no data-path collective",
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1409.2), d.get("hbm_gbs", 6569.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
_BEST_THREADS = None


def best_cpu_threads(frames) -> int:
    """torch's CPU kernels stop scaling (and then regress badly) well below the core count of a 100+-core host, so the CPU arm
    is timed with the fastest intra-op thread count among {8, 16, 32, 64, all} on one SuperPoint detection — the reference's
    own layout is a handful of threads per worker (gtsfm/runner.py:153-155)."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    import torch

    from gtsfm_b200 import synthetic as syn
    from oracle import superpoint_ref

    cores = os.cpu_count() or 1
    sd = syn.superpoint_state_dict(0)
    best, best_t = cores, float("inf")
    for n in sorted({min(c, cores) for c in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(n)
        superpoint_ref.detect_and_describe(frames[0], sd, MAX_KP)
        t0 = time.perf_counter()
        superpoint_ref.detect_and_describe(frames[0], sd, MAX_KP)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    _BEST_THREADS = best
    return best


def cpu_sample(frames, cal, n_pairs: int, threads: int):
    """Bounded sample of the SAME workload on the CPU: 2 detections (one reused), n_pairs match+verify."""
    import torch

    from gtsfm_b200 import synthetic as syn
    from oracle import lightglue_ref, superpoint_ref, verifier_ref

    torch.set_num_threads(threads)
    sp_sd, lg_sd = syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "bench")
    t0 = time.perf_counter()
    fa = superpoint_ref.detect_and_describe(frames[0], sp_sd, MAX_KP)
    t_det = time.perf_counter() - t0
    t_match = t_ver = 0.0
    for j in range(n_pairs):
        fb = superpoint_ref.detect_and_describe(frames[1 + j], sp_sd, MAX_KP)
        t0 = time.perf_counter()
        m = lightglue_ref.lightglue_match(fa[0], fa[2], fb[0], fb[2], lg_sd)
        t_match += time.perf_counter() - t0
        t0 = time.perf_counter()
        verifier_ref.verify_cv2(fa[0].astype(np.float64), fb[0].astype(np.float64), m.astype(np.uint32), cal, cal, True, THR_PX)
        t_ver += time.perf_counter() - t0
    per_pair = t_det / LOOKAHEAD + (t_match + t_ver) / n_pairs  # one detection serves `lookahead` pairs
    return 1.0 / per_pair, {"detect_s_per_frame": t_det, "match_s_per_pair": t_match / n_pairs, "verify_s_per_pair": t_ver / n_pairs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gtsfm_b200 import synthetic as syn

    frames, cal = syn.synthetic_sequence(4)
    cores = best_cpu_threads(frames)
    vals, stages = [], {}
    for i in range(args.warmup_ref + args.steps):
        v, stages = cpu_sample(frames, cal, 1, cores)
        if i >= args.warmup_ref:
            vals.append(v)
    value = float(np.mean(vals))
    sample = "per step: 1 SuperPoint detection + 1 LightGlue pair (5000x5000 keypoints, 9 layers) + 1 cv2 USAC verification; " \
             "pairs/s = 1 / (t_detect/20 + t_match + t_verify)"
    line = {
        "impl": "reference", "metric": "image_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup_ref, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port", "sample": sample, "stages": stages},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist

    from gtsfm_b200 import _lib, synthetic as syn, weights
    from gtsfm_b200.detector_descriptor import B200SuperPointDetectorDescriptor
    from gtsfm_b200.gtsfm_api import Cal3Bundler, Image
    from gtsfm_b200.matcher import B200LightGlueMatcher
    from gtsfm_b200.pipeline import DeviceFrontEnd
    from gtsfm_b200.verifier import B200Ransac

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # weights: rank 0 materialises them, one NCCL broadcast at start-up
    sp_sd, lg_sd = syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "bench")
    if world > 1:
        for sd, order in ((sp_sd, weights.SUPERPOINT_ORDER), (lg_sd, weights.LIGHTGLUE_ORDER)):
            blob = torch.from_numpy(np.concatenate([np.asarray(sd[k], np.float32).ravel() for k in order])).to(dev)
            dist.broadcast(blob, 0)
            flat, off = blob.cpu().numpy(), 0
            for k in order:
                n = int(np.prod(sd[k].shape)) if sd[k].shape else 1
                sd[k] = flat[off:off + n].reshape(sd[k].shape)
                off += n
    fe = DeviceFrontEnd(sp_sd, lg_sd, device=local, max_keypoints=MAX_KP)
    n_frames = LOOKAHEAD + NEW_FRAMES * (args.warmup + args.steps) * 2 + 4
    # each rank works on its own stretch of the sequence (weak scaling): different seed per rank
    frames, cal = syn.synthetic_sequence(n_frames, H, W, seed=77 + rank)
    frames_dev = [torch.from_numpy(f).to(dev) for f in frames]
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    window = deque(maxlen=LOOKAHEAD)
    for i in range(LOOKAHEAD):
        window.append(fe.detect(frames_dev[i]))
    cursor = LOOKAHEAD
    stats = {"matches": 0, "inliers": 0, "pairs": 0}

    def step_device(c):
        pending = []
        for j in range(NEW_FRAMES):
            f = fe.detect(frames_dev[c + j])
            prevs = list(window)
            for b0 in range(0, len(prevs), MATCH_BATCH):  # lock-step batches of 8 pairs (b2_lightglue_match_batched_dev)
                chunk = prevs[b0:b0 + MATCH_BATCH]
                for prev, (m, _) in zip(chunk, fe.match_batch([(prev, f) for prev in chunk])):
                    pending.append(fe.verify_async(prev, f, m, cal, cal, THR_PX))  # overlaps the next batch's matcher kernels
                    stats["matches"] += int(m.shape[0])
                    stats["pairs"] += 1
            window.append(f)
        for fut in pending:  # every verification result is collected inside the step
            stats["inliers"] += fut.result()[3]

    for _ in range(args.warmup):
        step_device(cursor)
        cursor += NEW_FRAMES
    stats.update(matches=0, inliers=0, pairs=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = fe.ctx.launch_count()
    vlaunch0 = fe._vctx.launch_count() if fe._vctx else 0
    fe.ctx.profile_start(DOMINANT_KERNEL)
    total_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush, outside the timed span
        torch.cuda.synchronize()
        e0.record()
        step_device(cursor)
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        cursor += NEW_FRAMES
    k_ms, k_launches, k_flop = fe.ctx.profile_stop()
    launches = fe.ctx.launch_count() - launches0 + (fe._vctx.launch_count() - vlaunch0 if fe._vctx else 0)
    # secondary figures (untimed region): detect-only rate (BASELINE configs[1] shape) and the encoder convolutions' rate
    torch.cuda.synchronize()
    fe.ctx.profile_start("k_conv_tma")
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for j in range(16):
        fe.detect(frames_dev[j % len(frames_dev)])
    d1.record()
    torch.cuda.synchronize()
    conv_ms, conv_n, conv_flop = fe.ctx.profile_stop()
    detect_ips = 16.0 / (d0.elapsed_time(d1) / 1e3)
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    pairs_total = PAIRS_PER_STEP * args.steps * world
    value = pairs_total / (max_ms / 1000.0)

    # ---- e2e: the same step through the GTSfM plugin API with host buffers ------------------------------------------
    det = B200SuperPointDetectorDescriptor(max_keypoints=MAX_KP, weights_path=sp_sd, device=local)
    mat = B200LightGlueMatcher("superpoint", weights_path=lg_sd, device=local)
    ver = B200Ransac(True, THR_PX, device=local)
    calib = Cal3Bundler(cal[0], 0, 0, cal[1], cal[2])
    hwin = deque(maxlen=LOOKAHEAD)
    for i in range(LOOKAHEAD):
        hwin.append(det.detect_and_describe(Image(frames[i])))
    def step_host(c):
        for j in range(NEW_FRAMES):
            kps, desc = det.detect_and_describe(Image(frames[c + j]))
            for pk, pd in list(hwin):
                m = mat.match(pk, kps, pd, desc, (H, W, 3), (H, W, 3))
                ver.verify(pk, kps, m, calib, calib)
            hwin.append((kps, desc))

    def copied():
        engs = [det._engine, mat._engine, ver._engine]
        return sum(e.h2d_bytes for e in engs if e), sum(e.d2h_bytes for e in engs if e)

    c2 = LOOKAHEAD
    for _ in range(min(args.warmup, 3)):
        step_host(c2)
        c2 += NEW_FRAMES
    barrier()
    h2d0, d2h0 = copied()
    e2e_wall = 0.0
    for i in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host(c2)
        torch.cuda.synchronize()
        e2e_wall += time.perf_counter() - t0
        c2 += NEW_FRAMES
    barrier()
    t = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total / float(t.item())
    h2d1, d2h1 = copied()
    traffic = {"h2d": (h2d1 - h2d0) // args.steps, "d2h": (d2h1 - d2h0) // args.steps}

    if rank == 0:
        tf_peak, hbm_peak, peak_src = measured_peaks()
        achieved = (k_flop / 1e12) / (k_ms / 1e3) if k_ms > 0 else 0.0
        line = {
            "metric": "image_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": CONFIG, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": traffic["h2d"], "d2h_bytes_per_step": traffic["d2h"]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": DOMINANT_KERNEL, "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak, "traffic": DOMINANT_KERNEL_DRAM_BYTES_PER_LAUNCH,
                         "traffic_unit": "bytes per launch (ncu, profiles/r01_flash_ps.txt)", "peak_source": f"bf16_tflops_sustained ({peak_src})",
                         "kernel_ms_per_step": k_ms / args.steps, "kernel_launches_per_step": k_launches / args.steps,
                         "kernel_share_of_step": k_ms / total_ms if total_ms else None},
            "work": {"matches_per_pair": stats["matches"] / max(1, stats["pairs"]), "inliers_per_pair": stats["inliers"] / max(1, stats["pairs"])},
            "extra": {"superpoint_detect_describe_images_per_sec_1gpu": detect_ips,
                      "encoder_conv_tflops": (conv_flop / 1e12) / (conv_ms / 1e3) if conv_ms > 0 else None,
                      "encoder_conv_frac_of_measured_bf16": ((conv_flop / 1e12) / (conv_ms / 1e3)) / tf_peak if conv_ms > 0 else None,
                      "note": "split-fp16 x3 products: tensor-pipe FLOPs are 3x the algorithmic FLOPs reported here"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = best_cpu_threads(frames)
            v, stages = cpu_sample(frames, cal, 2, cores)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                                    "sample": "oracle port (torch-CPU fp32 + cv2 USAC): 3 detections, 2 LightGlue pairs at 5000x5000 keypoints / 9 layers, "
                                              "2 verifications; pairs/s = 1 / (t_detect/20 + t_match + t_verify)", "stages": stages}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup_ref = min(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
