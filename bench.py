#!/usr/bin/env python
"""bench.py -- KLT tracked keypoints/s on the reference's headline workload (BASELINE.json configs[1]):
KITTI-shaped 1241x376 monocular stream, 2000 keypoints/frame, 3-level pyramid LK + forward-backward check,
batch of 64 frames per step on one B200.  A step = build 64 pyramids (update! path) + fb_tracking! of 64 x 2000
keypoints (frame i -> frame i+1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU; every rank owns an independent sequence (weak scaling, no
data-path collective; one small all_gather of tracked-keypoint counts at the end).  Prints ONE JSON line on rank 0.
`--impl reference` times the CPU restatement of the reference path (oracle/, all host cores) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 376, 1241
LEVELS, WINDOW, ITERS = 3, 9, 30
N_FRAMES, N_PTS = 64, 2000
MAX_DIST = 1.0
WORKLOAD = "kitti_mono_1241x376_2000kp_L3_w9_fb_batch64"
UNIT = "tracked keypoints/s"
METRIC = "KLT tracked keypoints/sec + frames/sec at 1241x376"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------- synthetic workload
def make_workload(seed: int, n_frames: int):
    """n_frames+1 consecutive frames; keypoints[i] live on frame i and are tracked into frame i+1."""
    from slamklt import synth
    frames_u8, affs = synth.make_sequence(seed, n_frames + 1, H, W)
    return frames_u8, affs


def topup_keypoints(kp_list, n_pts, seed):
    """Exactly n_pts sub-pixel keypoints per frame: detected corners first, random in-bounds points after."""
    from slamklt import synth
    out = np.empty((len(kp_list), n_pts, 2))
    for i, kp in enumerate(kp_list):
        kp = kp.astype(np.float64)[:n_pts]
        if len(kp) < n_pts:
            kp = np.vstack([kp, synth.random_keypoints(seed * 1000 + i, n_pts - len(kp), H, W)])
        out[i] = kp + np.random.default_rng(seed * 77 + i).uniform(-0.5, 0.5, kp.shape)
    return np.clip(out, 1.0, [H, W])


# --------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                clk, mx = float(p[1]), float(p[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take everything we saw
            for ts, line in self.lines:
                p = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(p[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_stream_time(frames_f64, pts, n_pairs, threads):
    """Wall time the reference path needs for n_pairs frames of a stream using `threads` cores: per frame one
    update!(pyramid) and one fb_tracking!."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.set_threads(1)
    pyrs = [O.LKPyramid(frames_f64[i], LEVELS, mode="ctor") for i in range(n_pairs + 1)]  # allocation, untimed

    def build(i):
        pyrs[i].update(frames_f64[i])

    def track(i):
        _, st, _ = O.fb_tracking(pyrs[i], pyrs[i + 1], pts[i], iterations=ITERS, window_size=WINDOW,
                                 pyramid_levels=LEVELS, max_distance=MAX_DIST)
        return int(st.sum())

    pyrs[0].update(frames_f64[0])  # previous frame of the first pair: carried over, untimed (as in a running stream)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(build, range(1, n_pairs + 1)))
        good = list(ex.map(track, range(n_pairs)))
    return time.perf_counter() - t0, sum(good)


# --------------------------------------------------------------------------------------- main
def bind_near_gpu(index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so the page-locked frame buffers are allocated
    (first touch) in the memory the GPU's PCIe root reads fastest.  Only at N > 1, where ranks would otherwise float over both
    sockets; best effort -- any failure leaves the affinity alone.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        cpus = open(dev + "/local_cpulist").read().strip()
        node = open(dev + "/numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids and len(ids) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, ids)
            return f"numa node {node}, {len(ids)} cpus"
        return f"numa node {node}, affinity unchanged"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="frame pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        from slamklt import synth
        n_pairs = args.cpu_sample or N_FRAMES
        frames_u8, _ = make_workload(2000, n_pairs)
        f64 = synth.to_f64(frames_u8)
        from oracle import oracle as O
        e = O.Extractor(2376, 17, (11, 36), 35)
        kps = topup_keypoints([O.detect(e, f64[i], np.zeros((0, 2))) for i in range(n_pairs)], N_PTS, 2000)
        times = []
        for it in range(args.warmup + args.steps):
            t, good = cpu_stream_time(f64, kps, n_pairs, cores)
            if it >= args.warmup:
                times.append(t)
        ms = 1e3 * float(np.mean(times))
        val = n_pairs * N_PTS / (ms / 1e3)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "frames_per_s": n_pairs / (ms / 1e3),
                "config": {"workload": WORKLOAD, "frames_per_step": n_pairs, "keypoints_per_frame": N_PTS,
                           "pyramid_levels": LEVELS, "window_size": WINDOW, "iterations": ITERS, "max_distance": MAX_DIST,
                           "note": "each step = one pass over n_pairs frame pairs of the batch-64 workload"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n_pairs} frame pairs of the same workload, frames spread over all cores"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import slamklt
    from slamklt import synth

    binding = bind_near_gpu(local_rank) if world > 1 and os.environ.get("SLAMKLT_NO_BIND") is None else "not bound (single rank)"
    log(f"[rank {rank}] cpu binding: {binding}")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner to stdout when the first communicator is created; stdout carries exactly one JSON
        # line, so file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            t = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist_.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        dist = dist_

    ctx = slamklt.Context(local_rank)
    alg = slamklt.LucasKanade(iterations=ITERS, window_size=WINDOW, pyramid_levels=LEVELS)
    seed = 2000 + rank
    t_gen = time.time()
    frames_u8, affs = make_workload(seed, N_FRAMES)
    f64 = synth.to_f64(frames_u8)
    log(f"[rank {rank}] generated {len(frames_u8)} frames in {time.time() - t_gen:.1f}s")

    # forward batch A: frames 1..64 tracked from 0..63; backward batch B: frames 63..0 tracked from 64..1 (palindrome keeps
    # the stream continuous across steps, so slamklt_batch_rotate carries a meaningful previous frame)
    batch = slamklt.StreamBatch(ctx, H, W, LEVELS, N_FRAMES, N_PTS)
    ext = slamklt.Extractor(2376, 17, (11, 36), 35)  # ceil(2376/396) = 6 corners per cell
    packA = slamklt.PinnedArray((N_FRAMES, W, H), np.float64)
    packB = slamklt.PinnedArray((N_FRAMES, W, H), np.float64)
    packA.array[...] = np.transpose(f64[1:], (0, 2, 1))
    packB.array[...] = np.transpose(f64[:-1][::-1], (0, 2, 1))
    # keypoints: detected on the GPU on the source frame of every pair (outside any timed region)
    dummy = np.zeros((N_FRAMES, N_PTS, 2)) + 50.0
    batch.prime(f64[0])
    src_fwd = np.ascontiguousarray(np.transpose(f64[:-1], (0, 2, 1)))       # frames 0..63
    batch.upload(src_fwd, dummy)
    kpA = topup_keypoints(batch.detect(ext), N_PTS, seed)
    src_bwd = np.ascontiguousarray(np.transpose(f64[1:][::-1], (0, 2, 1)))  # frames 64..1
    batch.upload(src_bwd, dummy)
    kpB = topup_keypoints(batch.detect(ext), N_PTS, seed + 1)
    ptsA = slamklt.PinnedArray((N_FRAMES, N_PTS, 2), np.float64); ptsA.array[...] = kpA
    ptsB = slamklt.PinnedArray((N_FRAMES, N_PTS, 2), np.float64); ptsB.array[...] = kpB
    outp = slamklt.PinnedArray((N_FRAMES, N_PTS, 2), np.float64)
    outs = slamklt.PinnedArray((N_FRAMES, N_PTS), np.uint8)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    # ---------------- device-resident value: inputs already in HBM, K x (build 64 pyramids + track 64x2000).
    # Two batch objects alternate (double buffering, as a stream consumer would): the tracking kernel of one batch runs on
    # the library's side stream while the next batch's pyramids are built, every step still does its full work.
    batch2 = slamklt.StreamBatch(ctx, H, W, LEVELS, N_FRAMES, N_PTS)
    pair = [batch, batch2]
    for b in pair:
        b.prime(f64[0])
        b.upload(packA.array, ptsA.array)
    ctx.sync()
    for i in range(args.warmup):
        pair[i % 2].process(alg, MAX_DIST)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the timed region lasts only tens of milliseconds: keep the same work running (untimed) for ~0.4 s first, so that the
    # 100 ms clock samples are taken under this load, then time exactly `steps` steps
    t_load0 = time.time()
    i = 0
    while time.time() - t_load0 < 0.4:
        pair[i % 2].process(alg, MAX_DIST); i += 1
        if i % 8 == 0:
            ctx.sync()
    barrier()
    ctx.stats(reset=True)
    launches1 = ctx.stats()["kernel_launches"]
    t_wall0 = time.time()
    ctx.timer_start()
    for i in range(args.steps):
        pair[i % 2].process(alg, MAX_DIST)   # build 64 pyramids + track 64 x 2000 keypoints (one C call)
    dev_ms = ctx.timer_stop()
    t_wall1 = time.time()
    barrier()
    clocks = sampler.stop(t_load0 + 0.1, t_wall1)
    clocks["window"] = "samples every 100 ms from 0.3 s of identical untimed load before the timed region to its end"
    st = ctx.stats()
    gpu_launches = st["kernel_launches"] - launches1
    lk_wpx, lk_it = st["lk_window_iters"], st["lk_iters"]
    _, status = batch.download()
    tracked_ok = int((status & 1).sum())

    # ---------------- phase split + per-kernel table (separate passes, not part of the timed region above)
    ctx.timer_start()
    for _ in range(args.steps):
        batch.build()
    build_ms = ctx.timer_stop() / args.steps
    ctx.timer_start()
    for _ in range(args.steps):
        batch.track(alg, MAX_DIST)
    track_ms = ctx.timer_stop() / args.steps
    ctx.profile(True)
    for _ in range(3):
        batch.build(); batch.track(alg, MAX_DIST)
    prof = ctx.profile_report()
    ctx.profile(False)
    kernels = {k: {"launches_per_step": n / 3, "ms_per_launch": ms / n} for k, (n, ms) in prof.items()}

    # ---------------- e2e: host buffers through the C ABI step call (H2D of 64 f64 frames + points, D2H of results)
    batch.prime(f64[0])
    h2d0 = ctx.stats()["h2d_bytes"]; d2h0 = ctx.stats()["d2h_bytes"]
    seq = [(packA, ptsA), (packB, ptsB)]
    for i in range(2):
        batch.step(seq[i % 2][0].array, seq[i % 2][1].array, alg, MAX_DIST, out_pts=outp.array, status=outs.array)
    barrier()
    s0 = ctx.stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        batch.step(seq[i % 2][0].array, seq[i % 2][1].array, alg, MAX_DIST, out_pts=outp.array, status=outs.array)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    s1 = ctx.stats()
    e2e_ok = int((outs.array & 1).sum())

    # same through u8 host frames (what a PNG decoder hands over; 8x fewer PCIe bytes), reported as extra information
    pack8A = slamklt.PinnedArray((N_FRAMES, W, H), np.uint8); pack8A.array[...] = np.transpose(frames_u8[1:], (0, 2, 1))
    pack8B = slamklt.PinnedArray((N_FRAMES, W, H), np.uint8); pack8B.array[...] = np.transpose(frames_u8[:-1][::-1], (0, 2, 1))
    batch.prime(f64[0])
    seq8 = [(pack8A, ptsA), (pack8B, ptsB)]
    for i in range(2):
        batch.step(seq8[i % 2][0].array, seq8[i % 2][1].array, alg, MAX_DIST, out_pts=outp.array, status=outs.array)
    ctx.sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        batch.step(seq8[i % 2][0].array, seq8[i % 2][1].array, alg, MAX_DIST, out_pts=outp.array, status=outs.array)
    ctx.sync()
    e2e8_s = time.perf_counter() - t0

    # ---------------- per-frame drop-in path (config[0] shape: one frame at a time through the reference-facing calls)
    single = None
    if world == 1:
        pa, pb = slamklt.LKPyramid(ctx, f64[0], LEVELS), slamklt.LKPyramid(ctx, f64[1], LEVELS)
        kp1k = kpA[0][:1000]
        ext1k = slamklt.Extractor(1000, 17, (11, 36), 35)
        # images in the layout the Julia caller hands over (column-major), so no host-side transpose is inside the timings
        fcol = [np.asfortranarray(f64[k]) for k in range(3)]
        pin = slamklt.PinnedArray((W, H), np.float64)          # the same frame in page-locked memory (slamklt_host_alloc)
        pin.array[...] = f64[1].T
        pin_img = pin.array.T                                   # (H, W) view, column-major
        u8col = np.asfortranarray(frames_u8[1])
        for i in range(3):
            pb.update(fcol[1 + i % 2]); slamklt.fb_tracking(pa, pb, kp1k, window_size=WINDOW, pyramid_levels=LEVELS, max_distance=MAX_DIST)
        t_upd, t_trk, t_det, t_pin, t_u8 = [], [], [], [], []
        for i in range(20):
            t0 = time.perf_counter(); pb.update(fcol[1 + i % 2]); t1 = time.perf_counter()
            r = slamklt.fb_tracking(pa, pb, kp1k, window_size=WINDOW, pyramid_levels=LEVELS, max_distance=MAX_DIST); t2 = time.perf_counter()
            slamklt.detect(ctx, ext1k, fcol[1], r[0][r[1]]); t3 = time.perf_counter()
            pb.update(pin_img); t4 = time.perf_counter()
            pb.update(u8col); t5 = time.perf_counter()
            t_upd.append(t1 - t0); t_trk.append(t2 - t1); t_det.append(t3 - t2); t_pin.append(t4 - t3); t_u8.append(t5 - t4)
        pb.update(fcol[1])
        pin.free()
        # optical_flow_matching! as one device call: 2000 keypoints, half of them 3-D with a projected prior
        kp2k = kpA[0][:N_PTS]
        sc = synth.matching_scene(5, kp2k, synth.true_flow(affs, 0, 1, kp2k))
        cam = slamklt.Camera(**sc["camera"])
        t_mat = []
        for i in range(23):
            t0 = time.perf_counter()
            slamklt.optical_flow_matching_frame(pa, pb, kp2k, sc["is_3d"], sc["world"], sc["cw"], cam, window_size=WINDOW,
                                                pyramid_levels=LEVELS, max_distance=MAX_DIST)
            if i >= 3:
                t_mat.append(time.perf_counter() - t0)
        single = {"update_ms": 1e3 * float(np.median(t_upd)), "fb_tracking_1000kp_ms": 1e3 * float(np.median(t_trk)),
                  "detect_ms": 1e3 * float(np.median(t_det)), "optical_flow_matching_2000kp_ms": 1e3 * float(np.median(t_mat)),
                  "update_pinned_f64_ms": 1e3 * float(np.median(t_pin)), "update_u8_ms": 1e3 * float(np.median(t_u8)),
                  "note": "host wall clock per call, Float64 host image in, results out (synchronous C ABI calls)"}

    # ---------------- max over ranks, gather of tracked-keypoint counts
    t_dev = dev_ms / 1e3
    t_e2e, t_e2e8 = e2e_s, e2e8_s
    from slam_jl_b200 import dist as skd
    dev = None
    if dist is not None:
        import torch
        dev = torch.device("cuda", local_rank)
    t_dev, t_e2e, t_e2e8 = skd.max_over_ranks([t_dev, t_e2e, t_e2e8], dist, dev)   # device time: MAX over ranks
    counts = skd.gather_counts(tracked_ok, dist, dev)  # the only collective: results stay with the rank that owns the sequence

    pts_per_step = world * N_FRAMES * N_PTS
    value = pts_per_step * args.steps / t_dev
    e2e_val = pts_per_step * args.steps / t_e2e
    e2e8_val = pts_per_step * args.steps / t_e2e8

    # ---------------- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    P0 = H * W
    dims = [(H, W)]
    for _ in range(LEVELS):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
    P = sum(h * w for h, w in dims)
    PL = dims[-1][0] * dims[-1][1]
    # B_pyr = b_in*P0 + 32*P - 4*P_L ; B_lk (no SATs) = min(sparse, dense)
    b_pyr = 8 * P0 + 32 * P - 4 * PL
    win = 2 * WINDOW + 1
    sparse = N_PTS * (LEVELS + 2) * (win * win * 3 * 4 + win * win * 12 + (win + 1) ** 2 * 4)
    dense = 28 * P + 28 * P0
    b_lk = min(sparse, dense)
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_launch"] * kv[1]["launches_per_step"]) if kernels else ("k_lk_fb", {"ms_per_launch": track_ms})
    dom_name, dom_ms = dom[0], dom[1]["ms_per_launch"]
    if dom_name.startswith("k_lk"):
        alg_bytes = b_lk * N_FRAMES
        alg_note = "B_lk=min(sparse,dense) per frame pair (no-SAT variant) x 64"
    else:
        lvl = int(dom_name.rsplit("L", 1)[1]) if "_L" in dom_name else 0
        px = dims[lvl][0] * dims[lvl][1] * N_FRAMES
        per_px = {"k_cols_all": 36 if lvl == 0 else 32, "k_rows_struct": 24, "k_rows_blur": 8, "k_resize": 5, "k_convert": 12}
        key = dom_name.rsplit("_L", 1)[0]
        alg_bytes = per_px.get(key, 8) * px
        alg_note = f"{per_px.get(key, 8)} B/px (reads+writes of that stage) x level-{lvl} pixels x 64"
    achieved = alg_bytes / (dom_ms / 1e3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "round1_traffic.json")))["dram_bytes_per_launch"].get(dom_name)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_note": alg_note, "ms_per_launch": dom_ms, "peak_source": peak_src}
    # whole-step view: all algorithmic bytes of a step over the step time
    step_bytes = (b_pyr + b_lk) * N_FRAMES
    ms_per_step = 1e3 * t_dev / args.steps
    step_roof = {"algorithmic_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms_per_step / 1e3) / 1e9,
                 "frac_of_hbm_peak": step_bytes / (ms_per_step / 1e3) / 1e9 / hbm_peak, "build_ms": build_ms, "track_ms": track_ms,
                 "pyramid_build_GBps": b_pyr * N_FRAMES / (build_ms / 1e3) / 1e9,
                 "pyramid_build_frac": b_pyr * N_FRAMES / (build_ms / 1e3) / 1e9 / hbm_peak}
    # LK against FP32 issue rate: 12 flop per window pixel per iteration
    lk_flops = 12.0 * lk_wpx / args.steps
    lk_roof = {"flop_per_step": lk_flops, "achieved_TFLOPs": lk_flops / (track_ms / 1e3) / 1e12,
               "avg_iterations_per_point_pass": lk_it / max(1, args.steps * N_FRAMES * N_PTS),
               "fp32_peak_TFLOPs_nominal": 148 * 128 * 2 * 1.965e9 / 1e12}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "dtype_note": "fp32 planes and window sums; Float64 2x2 solve, positions and decisions; Float64 extractor",
            "data": "synthetic", "frames_per_s": world * N_FRAMES * args.steps / t_dev,
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": N_FRAMES, "keypoints_per_frame": N_PTS,
                       "pyramid_levels": LEVELS, "window_size": WINDOW, "iterations": ITERS, "max_distance": MAX_DIST,
                       "buffers": "2 device-resident batches alternate (tracking of one overlaps the pyramid build of the other)",
                       "l2_policy": "working set per step (64 frames x 24.8 MB planes) is far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"{world} independent sequences, one per GPU" if world > 1 else "1 GPU",
                       "cpu_binding_rank0": binding},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps, "host_dtype": "f64",
                    "ms_per_step": 1e3 * t_e2e / args.steps, "tracked_ok_last_step": e2e_ok},
            "e2e_u8_host_frames": {"value": e2e8_val, "unit": UNIT, "ms_per_step": 1e3 * t_e2e8 / args.steps},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline, "step_roofline": step_roof, "lk_fp32": lk_roof, "kernels": kernels,
            "single_frame_calls": single, "tracked_ok_per_rank": counts, "tracked_fraction": tracked_ok / (N_FRAMES * N_PTS)}

    # ---------------- CPU baseline on rank 0, N = 1 only
    if rank == 0 and world == 1 and not args.no_cpu:
        n_pairs = args.cpu_sample or N_FRAMES
        t_cpu, good = min((cpu_stream_time(f64, kpA, n_pairs, cores) for _ in range(3)), key=lambda r: r[0])
        line["cpu_baseline"] = {"value": n_pairs * N_PTS / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_pairs} frame pairs of the same batch: update!(pyramid) + fb_tracking!, frames spread over "
                                          f"{cores} threads, best of 3 passes, {t_cpu:.2f}s wall per pass",
                                "frames_per_s": n_pairs / t_cpu, "tracked_ok": good}
        # the matching call of the single-frame leg on the CPU path (point loop threaded like lucas_kanade.jl:33)
        from oracle import oracle as O
        O.set_threads(cores)
        o0, o1 = O.LKPyramid(f64[0], LEVELS), O.LKPyramid(f64[1], LEVELS)
        o0.update(f64[0]); o1.update(f64[1])
        ocam = O.Camera(**sc["camera"])
        t_m = []
        for _ in range(3):
            t0 = time.perf_counter()
            O.optical_flow_matching(o0, o1, kp2k, sc["is_3d"], sc["world"], None, sc["cw"], ocam, window_size=WINDOW,
                                    pyramid_levels=LEVELS, max_distance=MAX_DIST)
            t_m.append(time.perf_counter() - t0)
        line["cpu_baseline"]["optical_flow_matching_2000kp_ms"] = 1e3 * min(t_m)
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line), flush=True)
    for p in (packA, packB, ptsA, ptsB, outp, outs, pack8A, pack8B):
        p.free()
    batch2.close()
    batch.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
