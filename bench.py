#!/usr/bin/env python
"""bench.py -- stereo frames/s of the B200-native front-end (BASELINE.json metric) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl ours|reference] [--workload b|full]

A "step" is one pass of the hot path over one batch of B synthetic stereo pairs.  At N=1 the workload is
BASELINE.json configs[1]: KITTI-shaped 1241x376 stereo, 1000 ORB/img, 8-level pyramid, L/R row-band matching.
  value      whole-job stereo frames/s with the batch already resident in HBM (device-timed, CUDA events on the
             library's stream, max over ranks).
  e2e        the same metric through the host-buffer C-ABI call: every step copies its images from pinned host
             memory to the device, runs all kernels and copies keypoints/descriptors/matches back.
  roofline   dominant kernel's achieved algorithmic GB/s (SURVEY.md 8(d) bytes) / MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (port of the reference's OpenCV-2.4/stereo-vo path) on a bounded sample.
N>1 (torchrun, one rank per GPU): frames shard across ranks with no data-path collective ("scaling": "weak").
`--impl reference` times the CPU oracle port on all host cores (the reference itself cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

W, H, NFEAT, NLEVELS = 1241, 376, 1000, 8
METRIC = "stereo frames/sec (1241x376, 1000 ORB, 8 levels, extract L+R + L/R row-band match)"
WORKLOAD = "configs[1]: 1241x376 stereo, 1000 ORB, 8 levels, L/R row-band match"
CONFIGS = {   # BASELINE.json configs; "b" is the headline (configs[1])
    "a": (1226, 370, 500, 1, "configs[0]: kitti.ini as shipped, 1226x370, 500 ORB, n_levels=1, L/R row-band match"),
    "b": (1241, 376, 1000, 8, WORKLOAD),
    "e": (1920, 1080, 4000, 8, "configs[4]: 1920x1080 stereo, 4000 ORB, 8 levels"),
}


def level_pixels(w=None, h=None, nlevels=None):
    w, h, nlevels = w or W, h or H, nlevels or NLEVELS
    # same arithmetic as cv::ORB (float 1.2^l, cvRound) -- sizes for the algorithmic-bytes table
    px = []
    for l in range(nlevels):
        s = np.float32(1.0) / np.float32(np.float64(np.float32(1.2)) ** l)
        px.append(int(np.rint(np.float32(w) * s)) * int(np.rint(np.float32(h) * s)))
    return px


def algorithmic_bytes_per_image(nkp=None):
    nkp = nkp or NFEAT
    """SURVEY.md 8(d): unfused-minimum HBM bytes per image for each stage."""
    px = level_pixels()
    P, A0 = sum(px), px[0]
    return {"pyramid": A0 + (P - A0), "fast_nms": P, "blur": 2 * P, "orient_brief": nkp * (37 * 37 + 60),
            "select_harris": 0, "match_rowband": 0, "P": P, "A0": A0}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []; self.p = None; self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its pinned staging buffers, first-touch) to the NUMA node the GPU hangs off:
    with one rank per GPU all ranks otherwise fight over one socket's memory controllers for their H2D traffic."""
    try:
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf                     # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def cpu_port_throughput(frames, threads, seconds_budget=20.0):
    """Oracle (CPU port of the reference path) on a bounded sample: extract L+R + row-band match per pair."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.lib()

    def one(pair):
        kl, dl, _ = orc.orb_extract(pair[0], nfeatures=NFEAT, nlevels=NLEVELS, fast_threshold=20)
        kr, dr, _ = orc.orb_extract(pair[1], nfeatures=NFEAT, nlevels=NLEVELS, fast_threshold=20)
        return len(orc.match_rowband(kl, dl, kr, dr, max_y_diff=2.0, max_dist=60))
    t0 = time.perf_counter(); one(frames[0]); t1 = time.perf_counter() - t0       # calibrate
    n = int(max(threads, min(len(frames) * 64, threads * max(1, int(seconds_budget / max(t1, 1e-3))))))
    work = [frames[i % len(frames)] for i in range(n)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:     # ctypes releases the GIL: real parallelism
        list(ex.map(one, work))
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from srba_stereo_slam_b200 import synth
    cores = os.cpu_count() or 1
    frames = synth.stereo_batch(4, W, H, seed=0, unique=4)
    vals = []
    for _ in range(args.warmup):
        cpu_port_throughput(frames, cores, seconds_budget=2.0)
    n_tot = 0; t_tot = 0.0
    for _ in range(args.steps):
        v, n, dt = cpu_port_throughput(frames, cores, seconds_budget=max(4.0, 60.0 / max(args.steps, 1)))
        n_tot += n; t_tot += dt
    value = n_tot / t_tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "stereo frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": n_tot // max(args.steps, 1)},
            "cpu_baseline": {"value": value, "unit": "stereo frames/s", "cores": cores, "kind": "port",
                             "sample": "%d stereo pairs over %d steps on %d host threads (oracle/frontend_oracle.c; the reference's OpenCV-2.4/stereo-vo path cannot be built here)" % (n_tot, args.steps, cores)},
            "e2e": {"value": value, "unit": "stereo frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=128, help="stereo pairs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="b", choices=sorted(CONFIGS), help="BASELINE.json workload (default: the headline configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the full front-end (BoW + KF match) measurement")
    args = ap.parse_args()
    global W, H, NFEAT, NLEVELS, METRIC, WORKLOAD
    W, H, NFEAT, NLEVELS, WORKLOAD = CONFIGS[args.config]
    METRIC = "stereo frames/sec (%dx%d, %d ORB, %d levels, extract L+R + L/R row-band match)" % (W, H, NFEAT, NLEVELS)
    if args.config == "e":
        args.pairs = min(args.pairs, 32)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local)
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from srba_stereo_slam_b200 import frontend, synth

    B = args.pairs
    frames = synth.stereo_batch(B, W, H, seed=rank, unique=8)        # [B,2,H,W] uint8, 8 rendered frames tiled
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=NLEVELS, nfeatures=NFEAT, max_pairs=B, device=local)
    stage = fe.host_staging()
    stage[:2 * B] = frames.reshape(-1, H, W)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local) if torch.cuda.is_available() else None
        fe.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    # Two contexts (each with its own stream and resident batch) take the steps alternately: the tail of one step's
    # kernels overlaps the head of the next, and issue-bound kernels (FAST) co-run with the shared-memory-bound ones
    # (describe) - measured +4-5 % over one context (tools/multi_ctx_probe.py).  One CUDA-event pair on the first
    # context's stream spans everything: sfe_stream_wait joins the second stream before the stop event.
    clocks = ClockSampler(local); clocks.start()        # spans warm-up, the resident loop and the end-to-end loop
    NCTX = 3
    extra = [frontend.StereoFrontEnd(width=W, height=H, nlevels=NLEVELS, nfeatures=NFEAT, max_pairs=B, device=local) for _ in range(NCTX - 1)]
    rfe = (fe, extra[0])
    for f in rfe:
        f.upload(frames)
    for i_ in range(2 * args.warmup):
        rfe[i_ % 2].process_device(B)
    for f in rfe:
        f.synchronize()
    kl, kr, dl, dr, m = fe.get_values(0)
    feats = (len(kl), len(kr), len(m))
    barrier()
    l0 = sum(f.kernel_launches() for f in rfe)
    fe.timer_start()
    for i_ in range(args.steps):
        rfe[i_ % 2].process_device(B)
    fe.stream_wait(rfe[1])
    ms = fe.timer_stop()
    rfe[1].synchronize()
    barrier()
    launches = sum(f.kernel_launches() for f in rfe) - l0
    if dist is not None:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H every step) -------------------------
    # Three contexts used round-robin (the documented streaming usage of sfe_submit_batch_host / sfe_wait_batch):
    # one batch's PCIe traffic overlaps the other's kernels.  Inputs live in pinned host memory; every step copies
    # its 2*B images to the device and its keypoints/descriptors/matches/counts back.
    fe2 = extra[0]
    for f in extra:
        f.host_staging()[:2 * B] = frames.reshape(-1, H, W)
    fes = (fe,) + tuple(extra)

    def sync_all():
        for f in fes:
            f.synchronize()
        barrier()
    for f in fes:
        f.submit_staged(B); f.wait_batch()
    sync_all()
    # link speed for context: one pinned->device copy of a step's images, timed alone
    t0 = time.perf_counter()
    for _ in range(5):
        fe.upload(stage[:2 * B].reshape(B, 2, H, W))
    h2d_gbs = 5 * 2 * B * W * H / (time.perf_counter() - t0) / 1e9
    sync_all()
    nmatch = 0
    t0 = time.perf_counter()
    for k_ in range(min(NCTX - 1, args.steps)):
        fes[k_].submit_staged(B)
    for s_ in range(args.steps):
        if s_ + NCTX - 1 < args.steps:
            fes[(s_ + NCTX - 1) % NCTX].submit_staged(B)
        ptrs = fes[s_ % NCTX].wait_batch()
        cnt = np.frombuffer((4 * 4 * B * __import__("ctypes").c_uint8).from_address(ptrs[3].value), np.int32).reshape(B, 4)
        nmatch += int(cnt[:, 2].sum())                      # the host reads the step's result
    sync_all()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop()
    if dist is not None:
        t = torch.tensor([ms_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_e2e = float(t.item())
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    cap = fe.kp_cap
    h2d = 2 * B * W * H
    d2h = 2 * B * cap * (28 + 32) + B * cap * 16 + 2 * B * 4 * 2 + B * 4

    # ---- full front-end (BASELINE.json configs[2] shape): + BoW query/add + keyframe store + association -----
    full = None
    voc_path = os.path.join(ROOT, "tests", "golden", "voc_flat.npz")
    if os.path.exists(voc_path) and not args.no_full:
        vz = np.load(voc_path)
        voc = {k: (vz[k] if vz[k].ndim else vz[k].item()) for k in vz.files}
        bow = frontend.BowDatabase(fe, voc)
        cap_frames = 4608 if B >= 64 else 36 * B                # >= KITTI-00's 4541 keyframes at the default batch
        kfdb = frontend.KeyframeDatabase(fe, bow, cap_frames)
        steps_full = min(args.steps, (cap_frames // B) - 4)
        prefill = (cap_frames // B) - steps_full - 1            # untimed batches so the timed ones see a large database
        # resident: two contexts with uploaded frames alternate, so the database work of one batch (latency-bound
        # gathers, on its source context's stream) overlaps the next batch's extraction (issue-bound) on the other stream
        rfe = [fe, fes[1]] if len(fes) > 1 and fes[1] is not fe else [fe]
        for f in rfe:
            f.upload(frames)
        for i_ in range(max(prefill, 3)):
            f = rfe[i_ % len(rfe)]
            f.process_device(B); kfdb.process_batch(B, src=f)
        for f in rfe:
            f.synchronize()
        db0 = kfdb.size()
        barrier()
        t0 = time.perf_counter()
        for i_ in range(steps_full):
            f = rfe[i_ % len(rfe)]
            f.process_device(B); kfdb.process_batch(B, src=f)
        for f in rfe:
            f.synchronize()
        ms_full = (time.perf_counter() - t0) * 1e3
        barrier()
        db1 = kfdb.size()
        if dist is not None:
            t = torch.tensor([ms_full], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_full = float(t.item())
        # end to end: H2D + everything + D2H of stereo results and of query hits / association keys
        kfdb.reset()
        nf = len(fes)
        for f in fes:
            f.submit_staged(B); kfdb.process_batch(B, src=f); f.wait_batch(); kfdb.results(B)
        steps_e = min(steps_full, (cap_frames - kfdb.size()) // B)
        barrier()
        for f in fes:
            f.synchronize()
        t0 = time.perf_counter()
        for k_ in range(min(nf - 1, steps_e)):
            fes[k_].submit_staged(B); kfdb.process_batch(B, src=fes[k_])
        hits = 0
        for s_ in range(steps_e):
            if s_ + nf - 1 < steps_e:
                nx = fes[(s_ + nf - 1) % nf]
                nx.submit_staged(B); kfdb.process_batch(B, src=nx)
            fes[s_ % nf].wait_batch()
            if s_ + 1 == steps_e:
                hits += int(kfdb.results(B)["query_n"].sum())      # results() syncs the database stream
        for f in fes:
            f.synchronize()
        barrier()
        ms_full_e2e = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([ms_full_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_full_e2e = float(t.item())
        full = {"metric": "stereo frames/sec, full front-end (extract + L/R match + BoW query/add + KF match vs previous + 4 hits)",
                "value": world * B * steps_full / (ms_full * 1e-3), "ms_per_step": ms_full / steps_full,
                "e2e": world * B * steps_e / (ms_full_e2e * 1e-3), "e2e_ms_per_step": ms_full_e2e / max(steps_e, 1),
                "db_entries_during_timing": [db0, db1], "steps": steps_full, "e2e_steps": steps_e,
                "note": "resident: 2 contexts alternate (database work overlaps the next extraction), host wall clock between full synchronisations; e2e: 3 contexts, downloads the association keys of the last batch only, stereo results every step"}
        kfdb.close(); bow.close()
    for f in extra:
        f.close()

    # ---- the reference's own call shape: one stereo pair per call (processNewImagePair + getValues), host buffers ----------
    fe1 = frontend.StereoFrontEnd(width=W, height=H, nlevels=NLEVELS, nfeatures=NFEAT, max_pairs=1, device=local)
    lat = []
    for i in range(25):
        t0 = time.perf_counter()
        fe1.process_stereo(frames[i % B, 0], frames[i % B, 1])
        fe1.get_values(0)
        lat.append((time.perf_counter() - t0) * 1e3)
    fe1.close()
    single = {"ms_per_pair_median": float(np.median(lat[5:])), "frames_per_s": 1e3 / float(np.median(lat[5:])),
              "note": "sfe_process_stereo + sfe_get_values on pageable host images, synchronous, batch of 1 (the live SLAM loop's shape)"}

    # ---- keyframe-to-keyframe matcher against the POPC pipe (BASELINE metric: "popc-pipe % of peak") --------------
    from srba_stereo_slam_b200 import loop_closure as lc
    NQ_M, K_M, NDB_M, NMAX_M = 512, 5, 2048, 500
    lcm = lc.LoopClosureMatcher(fe, NMAX_M, None)
    dbc, dbb = lc.synth_keyframe_blocks(NDB_M, NMAX_M, seed=100)
    rng_m = np.random.default_rng(7)
    src_m = rng_m.integers(0, NDB_M, NQ_M)
    qc, qb = lc.synth_keyframe_blocks(NQ_M, NMAX_M, seed=200, planted_from=(dbb, dbc, src_m))
    cand_m = rng_m.integers(0, NDB_M, (NQ_M, K_M)).astype(np.int32); cand_m[:, 0] = src_m
    lcm.set_database(dbc, dbb, NDB_M); lcm.upload_queries(qc, qb)
    pr = lcm.build_pairs(cand_m, "peer")
    for _ in range(3):
        lcm.run("peer")
    fe.synchronize(); fe.timer_start()
    for _ in range(10):
        lcm.run("peer")
    ms_m = fe.timer_stop() / 10
    dist_evals = float((pr["nq"].astype(np.int64) * pr["nt"].astype(np.int64)).sum())
    sm_ghz = (clk.get("sm_mhz") or 1900.0) / 1e3 if isinstance(clk, dict) else 1.9
    POPC_PEAK = 148 * 16 * sm_ghz * 1e9 / 1e12    # measured: 16 POPC/clk/SM (profiles/probes/popc_throughput_probe.cu) x 148 SMs x SM clock under load
    matcher = {"bound": "popc pipe", "kernel": "bf_pairs_kernel", "ms_per_launch": ms_m,
               "pairs": int(NQ_M * K_M), "distance_evaluations": dist_evals,
               "achieved": dist_evals * 8 / (ms_m * 1e-3) / 1e12, "peak": POPC_PEAK, "unit": "T popc32-equivalents/s",
               "frac": dist_evals * 8 / (ms_m * 1e-3) / 1e12 / POPC_PEAK,
               "popc_instructions_per_distance": 4,
               "popc_pipe_busy_frac": dist_evals * 4 / (ms_m * 1e-3) / 1e12 / POPC_PEAK,
               "note": "256-bit Hamming = 8 popc32 when done naively; the kernel's carry-save adder tree issues 4 POPC + 14 LOP3 per "
                       "distance, so `frac` can exceed 1 while the pipe itself is `popc_pipe_busy_frac` busy",
               "workload": "%d query KFs x %d candidates vs %d-KF database, %d descriptors/KF (configs[3] shape, one GPU)" % (NQ_M, K_M, NDB_M, NMAX_M)}
    lcm.close()

    # ---- per-kernel times, live (CUDA events around each stage, direct launches) --------------------------------
    fe.set_profiling(True)
    acc = {}
    reps = max(3, min(args.steps, 10))
    for i in range(reps + 1):
        fe.process_device(B); fe.synchronize()
        st = fe.stage_times()
        if i == 0:
            continue
        for k, (t_ms, n) in st.items():
            acc.setdefault(k, [0.0, n]); acc[k][0] += t_ms
    fe.set_profiling(False)
    stages = {k: {"ms": v[0] / reps, "launches": v[1]} for k, v in acc.items() if k not in ("h2d", "d2h")}
    ab = algorithmic_bytes_per_image(feats[0])
    nimg = 2 * B
    for k in stages:
        stages[k]["algorithmic_bytes"] = ab.get(k, 0) * nimg
        stages[k]["GBps"] = stages[k]["algorithmic_bytes"] / (stages[k]["ms"] * 1e-3) / 1e9 if stages[k]["ms"] > 0 else 0.0
    hbm_stages = [k for k in ("pyramid", "fast_nms", "blur", "orient_brief") if k in stages]
    dom = max(hbm_stages, key=lambda k: stages[k]["ms"])
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured"
    except Exception:
        pass
    per_launch_bytes = stages[dom]["algorithmic_bytes"] / max(stages[dom]["launches"], 1)
    per_launch_ms = stages[dom]["ms"] / max(stages[dom]["launches"], 1)
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    try:    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture, scaled to this batch size
        if args.config != "b":
            raise KeyError("traffic was captured for the headline config only")
        with open(os.path.join(ROOT, "profiles", "r01k_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["traffic_bytes_per_launch"][dom] * B / tj["pairs_per_launch"] / max(stages[dom]["launches"], 1)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                "stages": stages}

    # ---- CPU baseline (rank 0, N=1 only, bounded sample) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, n, dt = cpu_port_throughput(frames[:4], 1, seconds_budget=12.0)
        cpu = {"value": v, "unit": "stereo frames/s", "cores": 1, "kind": "port",
               "sample": "%d stereo pairs of the same workload in %.1f s, single thread (oracle/frontend_oracle.c)" % (n, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "stereo frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": B,
                           "keypoints_left_right_matches_pair0": feats,
                           "l2": "per-step working set %.0f MB (pyramid + blurred arenas of %d images) exceeds the 126 MB L2" % (2 * nimg * sum(level_pixels()) * 1.07 / 1e6, nimg),
                           "parallelism": "frames sharded over %d GPU(s), no data-path collective" % world, "rank0_numa_binding": numa},
                "e2e": {"value": e2e, "unit": "stereo frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "matches_read_back": nmatch,
                        "h2d_link_GBps_measured": h2d_gbs,
                        "timing": "host wall clock between full synchronisations; 3 contexts pipeline PCIe against kernels"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "full_frontend": full, "single_pair_latency": single, "matcher_roofline": matcher}
        print(json.dumps(line))
    fe.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
