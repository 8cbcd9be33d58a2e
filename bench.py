#!/usr/bin/env python
"""bench.py -- KLT tracked keypoints/s on the reference's workloads (BASELINE.json `configs`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c5]

Default (the driver's line) is c2 = BASELINE.json configs[1], the configuration the metric is quoted on: KITTI-shaped 1241x376
monocular stream, 2000 keypoints/frame, 3-level pyramid LK + forward-backward check, batch of 64 frames per step on one
B200.  A step = one pass of the hot path over one batch: build the batch's pyramids (update! path) + fb_tracking! of every
frame pair [+ stereo left->right matching (c1, c3)] [+ detect on every frame (c5)].
    c1  one stereo pair, 1000 keypoints, L=3        (configs[0], the reference's CPU-runnable example/kitty case)
    c2  64-frame mono batch, 2000 keypoints, L=3    (configs[1])
    c3  32 stereo pairs: temporal + left->right, 3000 keypoints, L=4   (configs[2])
    c5  16 frames 1920x1080, 8000 keypoints, L=5, detect every frame   (configs[4]; configs[3] is c2 at N > 1)
N > 1 is launched by torchrun, one rank per GPU; every rank owns an independent sequence (weak scaling, no data-path
collective; one small all_gather of tracked-keypoint counts, and -- timed separately -- the NCCL gather of one batch's
tracks).  Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference path (oracle/,
all host cores) on a bounded sample of the same workload.
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

WINDOW, ITERS, MAX_DIST = 9, 30, 1.0
UNIT = "tracked keypoints/s"
METRIC = "KLT tracked keypoints/sec + frames/sec at 1241x376"

# frames = frames (mono) or stereo pairs per step and per GPU; cpu_pairs = bounded sample of the CPU legs
CONFIGS = {
    "c1": dict(workload="kitti_stereo_pair_1241x376_1000kp_L3_w9_fb", H=376, W=1241, levels=3, n_pts=1000, frames=1, stereo=True,
               detect=False, cpu_pairs=16),
    "c2": dict(workload="kitti_mono_1241x376_2000kp_L3_w9_fb_batch64", H=376, W=1241, levels=3, n_pts=2000, frames=64, stereo=False,
               detect=False, cpu_pairs=64),
    "c3": dict(workload="kitti_stereo_temporal_1241x376_3000kp_L4_w9_fb_batch32pairs", H=376, W=1241, levels=4, n_pts=3000, frames=32,
               stereo=True, detect=False, cpu_pairs=16),
    "c5": dict(workload="hd_mono_1920x1080_8000kp_L5_w9_fb_detect_every_frame_batch16", H=1080, W=1920, levels=5, n_pts=8000, frames=16,
               stereo=False, detect=True, cpu_pairs=8),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def config_dict(cfg, world):
    """The `config` object of the JSON line: identical keys and values for both arms (the driver compares them)."""
    return {"workload": cfg["workload"], "frames_per_step_per_gpu": cfg["frames"], "keypoints_per_frame": cfg["n_pts"],
            "pyramid_levels": cfg["levels"], "window_size": WINDOW, "iterations": ITERS, "max_distance": MAX_DIST,
            "stereo": cfg["stereo"], "detect_every_frame": cfg["detect"], "image": f"{cfg['W']}x{cfg['H']}",
            "l2_policy": "working set per step is far larger than the 126 MB L2; no flush needed",
            "parallelism": f"{world} independent sequences, one per GPU" if world > 1 else "1 GPU"}


def points_per_step(cfg):
    return cfg["frames"] * cfg["n_pts"] * (2 if cfg["stereo"] else 1)


def extractor_for(cfg, O_or_slamklt, max_points=None):
    cs = 35
    grid = (-(-cfg["H"] // cs), -(-cfg["W"] // cs))
    return O_or_slamklt.Extractor(max_points or cfg["n_pts"] + grid[0] * grid[1], 17, grid, cs)


# --------------------------------------------------------------------------------------- synthetic workload
def make_workload(cfg, seed: int, n_frames: int):
    """n_frames+1 consecutive left frames (+ their right views for the stereo configs)."""
    from slamklt import synth
    frames_u8, affs = synth.make_sequence(seed, n_frames + 1, cfg["H"], cfg["W"])
    right_u8 = synth.right_views(frames_u8, seed) if cfg["stereo"] else None
    return frames_u8, right_u8, affs


def topup_keypoints(cfg, kp_list, seed):
    """Exactly n_pts sub-pixel keypoints per frame: detected corners first, random in-bounds points after."""
    from slamklt import synth
    n_pts, H, W = cfg["n_pts"], cfg["H"], cfg["W"]
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
def cpu_stream_time(cfg, left_f64, right_f64, pts, n_pairs, threads):
    """Wall time the reference path needs for n_pairs frames (stereo: pairs) of a stream using `threads` cores: per frame one
    update!(pyramid) and one fb_tracking! [+ update!(right pyramid) and the left->right fb_tracking!] [+ detect]."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.set_threads(1)
    L = cfg["levels"]
    pyrs = [O.LKPyramid(left_f64[i], L, mode="ctor") for i in range(n_pairs + 1)]  # allocation, untimed
    rpyrs = [O.LKPyramid(right_f64[i], L, mode="ctor") for i in range(n_pairs + 1)] if cfg["stereo"] else None
    ext = extractor_for(cfg, O)

    def build(i):
        pyrs[i].update(left_f64[i])
        if rpyrs:
            rpyrs[i].update(right_f64[i])

    def track(i):
        _, st, _ = O.fb_tracking(pyrs[i], pyrs[i + 1], pts[i], iterations=ITERS, window_size=WINDOW, pyramid_levels=L, max_distance=MAX_DIST)
        good = int(st.sum())
        if rpyrs:
            _, st2, _ = O.fb_tracking(pyrs[i], rpyrs[i], pts[i], iterations=ITERS, window_size=WINDOW, pyramid_levels=L, max_distance=MAX_DIST)
            good += int(st2.sum())
        if cfg["detect"]:
            O.detect(ext, left_f64[i + 1], pts[i][: (3 * cfg["n_pts"]) // 4])
        return good

    pyrs[0].update(left_f64[0])  # previous frame of the first pair: carried over, untimed (as in a running stream)
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


def level_dims(cfg):
    dims = [(cfg["H"], cfg["W"])]
    for _ in range(cfg["levels"]):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
    return dims


def algorithmic_bytes(cfg, b_in=8):
    """B_pyr = b_in*P0 + 32*P - 4*P_L per image; B_lk (no-SAT variant) = min(sparse, dense) per tracked frame pair;
    B_det = 4*P0 + 16*(n_cur + n_out) per detected frame."""
    dims = level_dims(cfg)
    P0 = dims[0][0] * dims[0][1]
    P = sum(h * w for h, w in dims)
    PL = dims[-1][0] * dims[-1][1]
    b_pyr = b_in * P0 + 32 * P - 4 * PL
    win = 2 * WINDOW + 1
    sparse = cfg["n_pts"] * (cfg["levels"] + 2) * (win * win * 3 * 4 + win * win * 12 + (win + 1) ** 2 * 4)
    dense = 28 * P + 28 * P0
    b_det = 4 * P0 + 16 * 2 * cfg["n_pts"]
    return dict(P0=P0, P=P, dims=dims, b_pyr=b_pyr, b_lk=min(sparse, dense), b_det=b_det)


def run_reference(args, cfg, cores):
    """--impl reference: the CPU port of the reference path (oracle/), all host cores, bounded sample of the same workload."""
    from slamklt import synth
    from oracle import oracle as O
    n_pairs = args.cpu_sample or cfg["cpu_pairs"]
    left_u8, right_u8, _ = make_workload(cfg, 2000, n_pairs)
    lf = synth.to_f64(left_u8)
    rf = synth.to_f64(right_u8) if right_u8 is not None else None
    e = extractor_for(cfg, O)
    kps = topup_keypoints(cfg, [O.detect(e, lf[i], np.zeros((0, 2))) for i in range(n_pairs)], 2000)
    times = []
    for it in range(args.warmup + args.steps):
        t, good = cpu_stream_time(cfg, lf, rf, kps, n_pairs, cores)
        if it >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    ppp = points_per_step(cfg) // cfg["frames"]
    val = n_pairs * ppp / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "frames_per_s": n_pairs / (ms / 1e3),
            "config": config_dict(cfg, args.gpus),
            "sample_note": f"each step = one pass over {n_pairs} frame pairs of the workload (a full step has {cfg['frames']}), spread over all cores",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_pairs} frame pairs of the same workload, frames spread over all cores"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="frame pairs in the CPU baseline sample (0 = per-config default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development only)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    H, W, LEVELS, N_PTS, NF = cfg["H"], cfg["W"], cfg["levels"], cfg["n_pts"], cfg["frames"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank == 0:
            if args.steps > 20:  # the CPU leg is seconds per step: keep the whole run within minutes
                args.steps = 20
            run_reference(args, cfg, cores)
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

    # host threads the library may use to repack Float64 host frames (slamklt_batch_step): the ranks of one box share its cores
    if world > 1 and "SLAMKLT_HOST_THREADS" not in os.environ:
        os.environ["SLAMKLT_HOST_THREADS"] = str(max(1, min(16, cores // world)))
    ctx = slamklt.Context(local_rank)
    alg = slamklt.LucasKanade(iterations=ITERS, window_size=WINDOW, pyramid_levels=LEVELS)
    seed = 2000 + rank
    t_gen = time.time()
    left_u8, right_u8, affs = make_workload(cfg, seed, NF)
    f64 = synth.to_f64(left_u8)
    log(f"[rank {rank}] generated {len(left_u8)} frames in {time.time() - t_gen:.1f}s")

    def packed(frames, dtype):
        a = slamklt.PinnedArray((NF, W, H), dtype)
        a.array[...] = np.transpose(frames, (0, 2, 1))
        return a

    # forward sequence A: frames 1..NF tracked from 0..NF-1; reversed sequence B: frames NF-1..0 tracked from NF..1 (a palindrome
    # keeps the stream continuous across steps, so slamklt_batch_rotate carries a meaningful previous frame)
    batch = slamklt.StreamBatch(ctx, H, W, LEVELS, NF, N_PTS)
    ext = extractor_for(cfg, slamklt)
    packA, packB = packed(f64[1:], np.float64), packed(f64[:-1][::-1], np.float64)
    # keypoints: detected on the GPU on the source frame of every pair (outside any timed region)
    dummy = np.zeros((NF, N_PTS, 2)) + 50.0
    batch.prime(f64[0])
    batch.upload(np.ascontiguousarray(np.transpose(f64[:-1], (0, 2, 1))), dummy)        # frames 0..NF-1
    kpA = topup_keypoints(cfg, batch.detect(ext), seed)
    batch.upload(np.ascontiguousarray(np.transpose(f64[1:][::-1], (0, 2, 1))), dummy)   # frames NF..1
    kpB = topup_keypoints(cfg, batch.detect(ext), seed + 1)
    ptsA = slamklt.PinnedArray((NF, N_PTS, 2), np.float64); ptsA.array[...] = kpA
    ptsB = slamklt.PinnedArray((NF, N_PTS, 2), np.float64); ptsB.array[...] = kpB
    outp = slamklt.PinnedArray((NF, N_PTS, 2), np.float64)
    outs = slamklt.PinnedArray((NF, N_PTS), np.uint8)
    pinned = [packA, packB, ptsA, ptsB, outp, outs]
    # stereo configs: the right frames live in a second batch whose slot i+1 pairs with the left batch's slot i+1
    rbatch = rpackA = None
    if cfg["stereo"]:
        r64 = synth.to_f64(right_u8)
        rbatch = slamklt.StreamBatch(ctx, H, W, LEVELS, NF, N_PTS)
        rbatch.prime(r64[0])
        # the stereo match of pair i tracks the keypoints of LEFT frame i+1 (the frame just built) into RIGHT frame i+1
        rpackA = packed(r64[1:], np.float64)
        rpts = slamklt.PinnedArray((NF, N_PTS, 2), np.float64)
        rpts.array[...] = np.concatenate([kpA[1:], kpA[-1:]], axis=0)
        routp = slamklt.PinnedArray((NF, N_PTS, 2), np.float64)
        routs = slamklt.PinnedArray((NF, N_PTS), np.uint8)
        pinned += [rpackA, rpts, routp, routs]
    det_cur = np.ascontiguousarray(kpA[:, : (3 * N_PTS) // 4]) if cfg["detect"] else None  # 3/4 of the keypoints survive; detect tops up

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    # ---------------- device-resident value: inputs already in HBM, K x one step.
    # Two batch objects alternate (double buffering, as a stream consumer would): the tracking kernel of one batch runs on
    # the library's side stream while the next batch's pyramids are built; every step still does its full work.
    batch2 = slamklt.StreamBatch(ctx, H, W, LEVELS, NF, N_PTS)
    pair = [batch, batch2]
    for b in pair:
        b.prime(f64[0])
        b.upload(packA.array, ptsA.array)
    if rbatch is not None:
        rbatch.upload(rpackA.array, rpts.array)
    ctx.sync()

    def device_step(i):
        b = pair[i % 2]
        b.process(alg, MAX_DIST)               # build NF pyramids + track NF x N_PTS keypoints (one C call)
        if rbatch is not None:
            rbatch.build()                     # right pyramids
            b.track_cross(rbatch, alg, MAX_DIST)   # left -> right matching of the same keypoints
        if cfg["detect"]:
            b.detect(ext, det_cur)             # re-extraction on every frame (synchronous: results come back to the host)

    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # keep the same work running (untimed) for ~0.4 s first, so that the 100 ms clock samples are taken under this load
    t_load0 = time.time()
    i = 0
    while time.time() - t_load0 < 0.4:
        device_step(i); i += 1
        if i % 8 == 0:
            ctx.sync()
    barrier()
    ctx.stats(reset=True)
    launches1 = ctx.stats()["kernel_launches"]
    ctx.timer_start()
    for i in range(args.steps):
        device_step(i)
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
    psteps = max(3, min(args.steps, 20))
    ctx.timer_start()
    for _ in range(psteps):
        batch.build()
    build_ms = ctx.timer_stop() / psteps
    ctx.timer_start()
    for _ in range(psteps):
        batch.track(alg, MAX_DIST)
    track_ms = ctx.timer_stop() / psteps
    detect_ms = None
    if cfg["detect"]:
        t0 = time.perf_counter()
        for _ in range(psteps):
            batch.detect(ext, det_cur)
        detect_ms = 1e3 * (time.perf_counter() - t0) / psteps
    ctx.profile(True)
    for _ in range(3):
        batch.build(); batch.track(alg, MAX_DIST)
        if cfg["detect"]:
            batch.detect(ext, det_cur)
    prof = ctx.profile_report()
    ctx.profile(False)
    kernels = {k: {"launches_per_step": n / 3, "ms_per_launch": ms / n} for k, (n, ms) in prof.items()}

    # ---------------- e2e: host buffers through the C ABI (H2D of the step's frames + points, D2H of its results, every step)
    outp2 = slamklt.PinnedArray((NF, N_PTS, 2), np.float64)
    outs2 = slamklt.PinnedArray((NF, N_PTS), np.uint8)
    pinned += [outp2, outs2]
    results = [(outp, outs), (outp2, outs2)]
    simple = rbatch is None   # mono configs go through slamklt_batch_step (stereo configs pair two batches call by call)

    def e2e_loop(seq, rseq, in_flight=2):
        """seq: [(frames, pts), (frames, pts)] alternating host buffers (palindrome); returns seconds for args.steps steps.
        in_flight = 2 (mono configs): two batches -- two independent streams, NF consecutive frames each per step -- alternate, each step queued with
        slamklt_batch_step_begin and collected with slamklt_batch_step_end one step later, so the uploads of one step overlap
        the kernels of the other; every step still uploads its own NF frames + points and downloads its own results inside the
        timed region.  in_flight = 1: one synchronous slamklt_batch_step per step."""
        for b in pair:
            b.prime(f64[0])

        def one(i):
            fr, pt = seq[i % 2]
            if simple:
                batch.step(fr.array, pt.array, alg, MAX_DIST, out_pts=outp.array, status=outs.array)   # one pipelined C call
                if cfg["detect"]:
                    batch.detect(ext, det_cur)
                return
            batch.upload(fr.array, pt.array)
            batch.process(alg, MAX_DIST)
            if rbatch is not None:
                rbatch.upload(rseq[i % 2].array, rpts.array)
                rbatch.build()
                batch.track_cross(rbatch, alg, MAX_DIST)
                rbatch.download(routp.array, routs.array)
            batch.download(outp.array, outs.array)
            if cfg["detect"]:
                batch.detect(ext, det_cur)
            batch.rotate()

        def begin(i):
            fr, pt = seq[(i // 2) % 2]          # each batch walks its own palindrome
            o, s_ = results[i % 2]
            pair[i % 2].step_begin(fr.array, pt.array, alg, MAX_DIST, out_pts=o.array, status=s_.array)

        def run(n):
            if not (simple and in_flight == 2):
                for i in range(n):
                    one(i)
                return
            def end(i):
                pair[i % 2].step_end()
                if cfg["detect"]:
                    pair[i % 2].detect(ext, det_cur)    # re-extraction on the frames of the step just finished (synchronous)

            begin(0)
            for i in range(1, n):
                begin(i)
                end(i - 1)
            end(n - 1)

        run(4)
        barrier()
        s0 = ctx.stats()
        t0 = time.perf_counter()
        run(args.steps)
        ctx.sync()
        dt = time.perf_counter() - t0
        s1 = ctx.stats()
        return dt, (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps, (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps

    rpackB = packed(synth.to_f64(right_u8)[:-1][::-1], np.float64) if rbatch is not None else None
    if rpackB is not None:
        pinned.append(rpackB)
    e2e_s, e2e_h2d, e2e_d2h = e2e_loop([(packA, ptsA), (packB, ptsB)], [rpackA, rpackB])
    e2e_ok = int((outs.array & 1).sum())
    # for comparison: one synchronous slamklt_batch_step per step (nothing overlaps across steps)
    e2e1_s = e2e_loop([(packA, ptsA), (packB, ptsB)], [rpackA, rpackB], in_flight=1)[0] if simple else e2e_s
    # re-extraction after every step is a synchronous call that keeps the host from repacking the next step: with it, one step at a
    # time can be the faster schedule -- report the better one and say which
    e2e_in_flight = 2 if simple else 1
    if cfg["detect"] and e2e1_s < e2e_s:
        e2e_s, e2e_in_flight = e2e1_s, 1
    # the same through u8 host frames: what a camera / PNG decoder hands over (the reference's example converts to Gray{Float64}
    # on the host, example/kitty/main.jl:36-40); 8x fewer PCIe bytes, bit-identical pyramids (tests/test_gpu_batch.py)
    pack8A, pack8B = packed(left_u8[1:], np.uint8), packed(left_u8[:-1][::-1], np.uint8)
    pinned += [pack8A, pack8B]
    r8A = r8B = None
    if rbatch is not None:
        r8A, r8B = packed(right_u8[1:], np.uint8), packed(right_u8[:-1][::-1], np.uint8)
        pinned += [r8A, r8B]
    e2e8_s, e2e8_h2d, e2e8_d2h = e2e_loop([(pack8A, ptsA), (pack8B, ptsB)], [r8A, r8B])
    e2e81_s = e2e_loop([(pack8A, ptsA), (pack8B, ptsB)], [r8A, r8B], in_flight=1)[0] if simple else e2e8_s
    e2e8_in_flight = 2 if simple else 1
    if cfg["detect"] and e2e81_s < e2e8_s:
        e2e8_s, e2e8_in_flight = e2e81_s, 1
    up = ctx.upload_rates()

    # ---------------- per-frame drop-in path (one frame at a time through the reference-facing calls), c2 at N = 1 only
    single = None
    sc = kp2k = None
    if world == 1 and args.config == "c2":
        pa, pb = slamklt.LKPyramid(ctx, f64[0], LEVELS), slamklt.LKPyramid(ctx, f64[1], LEVELS)
        kp1k = kpA[0][:1000]
        ext1k = slamklt.Extractor(1000, 17, (11, 36), 35)
        # images in the layout the Julia caller hands over (column-major), so no host-side transpose is inside the timings
        fcol = [np.asfortranarray(f64[k]) for k in range(3)]
        pin = slamklt.PinnedArray((W, H), np.float64)          # the same frame in page-locked memory (slamklt_host_alloc)
        pin.array[...] = f64[1].T
        pin_img = pin.array.T                                   # (H, W) view, column-major
        u8col = np.asfortranarray(left_u8[1])
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

    # ---------------- max over ranks, gather of tracked-keypoint counts, NCCL gather of one batch's tracks (north_star)
    t_dev = dev_ms / 1e3
    from slam_jl_b200 import dist as skd
    dev = None
    gather_us = None
    if dist is not None:
        import torch
        dev = torch.device("cuda", local_rank)
        # the (NF x N_PTS x 2 f64 + NF x N_PTS u8) result gather a single consumer would ask for: outside the device-timed region
        p_dev = torch.as_tensor(outp.array, device=dev)
        s_dev = torch.as_tensor(outs.array, device=dev)
        for _ in range(3):
            skd.gather_tracks(p_dev, s_dev, dist, dev)
        torch.cuda.synchronize(); dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            ps, ss = skd.gather_tracks(p_dev, s_dev, dist, dev)
        ev1.record(); torch.cuda.synchronize()
        gather_us = 1e3 * ev0.elapsed_time(ev1) / 10
        # (positions of failed tracks are NaN: compare bit patterns)
        assert len(ps) == world and torch.equal(ps[rank].view(torch.int64), p_dev.view(torch.int64)) and torch.equal(ss[rank], s_dev)
    t_dev, t_e2e, t_e2e8, t_e2e1, t_e2e81 = skd.max_over_ranks([t_dev, e2e_s, e2e8_s, e2e1_s, e2e81_s], dist, dev)   # MAX over ranks
    if gather_us is not None:
        gather_us = skd.max_over_ranks([gather_us], dist, dev)[0]
    counts = skd.gather_counts(tracked_ok, dist, dev)  # results stay with the rank that owns the sequence

    pts_per_step = world * points_per_step(cfg)
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
    ab = algorithmic_bytes(cfg)
    dims = ab["dims"]
    n_img = NF * (2 if cfg["stereo"] else 1)
    n_lk = NF * (2 if cfg["stereo"] else 1)
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_launch"] * kv[1]["launches_per_step"]) if kernels else ("k_lk_fb", {"ms_per_launch": track_ms})
    dom_name, dom_ms = dom[0], dom[1]["ms_per_launch"]
    if dom_name.startswith("k_lk"):
        alg_bytes = ab["b_lk"] * NF
        alg_note = f"B_lk=min(sparse,dense) per frame pair (no-SAT variant) x {NF}"
    elif dom_name.startswith("k_detect"):
        alg_bytes = ab["b_det"] * NF
        alg_note = f"B_det = 4*P0 + 16*(n_cur + n_out) per frame x {NF}"
    else:
        lvl = int(dom_name.rsplit("L", 1)[1]) if "_L" in dom_name else 0
        px = dims[lvl][0] * dims[lvl][1] * NF
        per_px = {"k_cols_all": 36 if lvl == 0 else 32, "k_cols_grad": 28, "k_cols_blur": 8, "k_rows_struct": 24, "k_rows_blur": 8, "k_resize": 5, "k_convert": 12}
        key = dom_name.rsplit("_L", 1)[0]
        alg_bytes = per_px.get(key, 8) * px
        alg_note = f"{per_px.get(key, 8)} B/px (reads+writes of that stage) x level-{lvl} pixels x {NF}"
    achieved = alg_bytes / (dom_ms / 1e3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (c2 workload)
        if args.config == "c2":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "round2_traffic.json")))["dram_bytes_per_launch"].get(dom_name)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_note": alg_note, "ms_per_launch": dom_ms, "peak_source": peak_src,
                "note": "the tracking kernel is bound by instruction issue, not by HBM: see lk_issue (SURVEY 8d reports LK against the FP32 issue rate)"
                        if dom_name.startswith("k_lk") else None}
    # whole-step view: all algorithmic bytes of a step over the step time
    step_bytes = ab["b_pyr"] * n_img + ab["b_lk"] * n_lk + (ab["b_det"] * NF if cfg["detect"] else 0)
    ms_per_step = 1e3 * t_dev / args.steps
    step_roof = {"algorithmic_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms_per_step / 1e3) / 1e9,
                 "frac_of_hbm_peak": step_bytes / (ms_per_step / 1e3) / 1e9 / hbm_peak, "build_ms": build_ms, "track_ms": track_ms,
                 "detect_ms_wall": detect_ms,
                 "pyramid_build_GBps": ab["b_pyr"] * NF / (build_ms / 1e3) / 1e9,
                 "pyramid_build_frac": ab["b_pyr"] * NF / (build_ms / 1e3) / 1e9 / hbm_peak}
    # LK against the FP32 issue rate: 12 flop per window pixel per iteration, plus the committed ncu counters of the
    # tracking kernel (instructions per keypoint, issue utilisation) -- the quantities that bound it
    lk_flops = 12.0 * lk_wpx / args.steps
    lk_passes = max(1, args.steps * n_lk * N_PTS)
    lk_issue = {"bound": "issue", "flop_per_step": lk_flops, "achieved_TFLOPs": lk_flops / (track_ms * (2 if cfg["stereo"] else 1) / 1e3) / 1e12,
                "avg_iterations_per_keypoint": lk_it / lk_passes,
                "fp32_peak_TFLOPs_nominal": 148 * 128 * 2 * 1.965e9 / 1e12, "track_ms_per_launch": track_ms,
                "keypoints_per_launch": NF * N_PTS}
    try:
        lk_issue["ncu"] = json.load(open(os.path.join(ROOT, "profiles", "round2_lk_issue.json")))
    except Exception:
        lk_issue["ncu"] = None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "dtype_note": "fp32 planes, window sums and in-level position fractions; Float64 structure-tensor inverse, point "
                                          "coordinates and gates; Float64 extractor",
            "data": "synthetic", "frames_per_s": world * NF * (2 if cfg["stereo"] else 1) * args.steps / t_dev,
            "config": config_dict(cfg, world),
            "run_notes": {"buffers": "2 device-resident batches alternate (tracking of one overlaps the pyramid build of the other)",
                          "cpu_binding_rank0": binding},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(e2e_h2d), "d2h_bytes_per_step": int(e2e_d2h), "host_dtype": "f64",
                    "ms_per_step": 1e3 * t_e2e / args.steps, "tracked_ok_last_step": e2e_ok,
                    "h2d_GBps_per_rank": e2e_h2d / (t_e2e / args.steps) / 1e9,
                    "host_threads": int(os.environ.get("SLAMKLT_HOST_THREADS", min(16, cores))),
                    "steps_in_flight": e2e_in_flight,
                    "ms_per_step_one_synchronous_call": 1e3 * t_e2e1 / args.steps,
                    "host_source_GBps_all_ranks": world * NF * (2 if cfg["stereo"] else 1) * H * W * 8 / (t_e2e / args.steps) / 1e9,
                    "host_source_note": "bytes of Float64 frames the ranks of this box read from host memory per second (by worker threads "
                                        "or by the copy engines): at N >= 2 this number, not the GPUs, bounds the Float64 e2e figure "
                                        "(measured: about 180 GB/s on the 32-vCPU 8-GPU host); UInt8 frames (e2e_u8) are 8x lighter and scale",
                    "upload_engines_rank0": up,
                    "note": "Float64 host frames (the reference's Matrix{Gray{Float64}}) in page-locked memory.  When every pixel is an exact "
                            "k/255 -- 8-bit camera data, as in the reference's example -- host worker threads repack some chunks of a step to 8 "
                            "bits (lossless, bit-identical pyramids) while the copy engine ships the other chunks as plain Float64: both "
                            "engines read host memory at once, the split follows their measured rates (upload_engines_rank0, bytes of "
                            "source per second).  Frames that are not 8-bit data travel as they are.  steps_in_flight = 2: two batches "
                            "(two independent streams) alternate through slamklt_batch_step_begin/_end, so the uploads of one "
                            "step overlap the kernels of the other; every step's uploads and result downloads are inside the timed region"},
            "e2e_u8": {"value": e2e8_val, "unit": UNIT, "h2d_bytes_per_step": int(e2e8_h2d), "d2h_bytes_per_step": int(e2e8_d2h),
                       "host_dtype": "u8", "ms_per_step": 1e3 * t_e2e8 / args.steps,
                       "steps_in_flight": e2e8_in_flight, "ms_per_step_one_synchronous_call": 1e3 * t_e2e81 / args.steps,
                       "h2d_GBps_per_rank": e2e8_h2d / (t_e2e8 / args.steps) / 1e9,
                       "note": "same call with UInt8 host frames (camera / PNG decoder output), converted on the device with the reference's "
                               "i/255 semantics; pyramids bit-identical to the Float64 upload"},
            "gather_us": gather_us,
            "gpu_launches": int(gpu_launches),
            "roofline": roofline, "step_roofline": step_roof, "lk_issue": lk_issue, "kernels": kernels,
            "single_frame_calls": single, "tracked_ok_per_rank": counts, "tracked_fraction": tracked_ok / (NF * N_PTS)}

    # ---------------- CPU baseline on rank 0, N = 1 only
    if rank == 0 and world == 1 and not args.no_cpu:
        n_pairs = min(args.cpu_sample or cfg["cpu_pairs"], NF)
        rf = synth.to_f64(right_u8) if right_u8 is not None else None
        t_cpu, good = min((cpu_stream_time(cfg, f64, rf, kpA, n_pairs, cores) for _ in range(2)), key=lambda r: r[0])
        ppp = points_per_step(cfg) // NF
        line["cpu_baseline"] = {"value": n_pairs * ppp / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_pairs} frame pairs of the same batch: update!(pyramid) + fb_tracking!"
                                          f"{' + right update! + left->right fb_tracking!' if cfg['stereo'] else ''}"
                                          f"{' + detect' if cfg['detect'] else ''}, frames spread over "
                                          f"{cores} threads, best of 2 passes, {t_cpu:.2f}s wall per pass",
                                "frames_per_s": n_pairs / t_cpu, "tracked_ok": good}
        if sc is not None:
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
    for p in pinned:
        p.free()
    batch2.close()
    if rbatch is not None:
        rbatch.close()
    batch.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
