#!/usr/bin/env python
"""bench.py -- stereo frames/s of the B200-native front-end (BASELINE.json metric) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl ours|reference] [--workload b|full]

A "step" is one pass of the hot path over --batches-per-step (default 8) batches of B = --pairs (default 128) synthetic stereo
pairs, i.e. 1024 pairs per GPU and step.  At N=1 the workload is BASELINE.json configs[1]: KITTI-shaped 1241x376 stereo,
1000 ORB/img, 8-level pyramid, L/R row-band matching.
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


def bench_config(pairs_per_step):
    """The `config` object of the JSON line -- identical keys and values in both arms (ours / --impl reference) for the same
    command line, so the driver's same_config check holds; everything arm-specific goes to `details`."""
    return {"workload": WORKLOAD, "width": W, "height": H, "orb_features": NFEAT, "pyramid_levels": NLEVELS, "fast_threshold": 20,
            "lr_match": "row band +-2 px, Hamming <= 60", "pairs_per_step_per_gpu": pairs_per_step}


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
    """--impl reference: the CPU port of the reference's path (the oracle; the reference's own OpenCV-2.4 / stereo-vo code cannot be
    built here) on ALL host threads, same metric / config; each step is a bounded sample of the step's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from srba_stereo_slam_b200 import synth
    cores = os.cpu_count() or 1
    frames = synth.stereo_batch(4, W, H, seed=0, unique=4)
    for _ in range(args.warmup):
        cpu_port_throughput(frames, cores, seconds_budget=2.0)
    n_tot = 0; t_tot = 0.0
    for _ in range(args.steps):
        v, n, dt = cpu_port_throughput(frames, cores, seconds_budget=max(4.0, 60.0 / max(args.steps, 1)))
        n_tot += n; t_tot += dt
    value = n_tot / t_tot
    pps = args.pairs * args.batches_per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "stereo frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(pps),
            "details": {"sampled_pairs_per_step": n_tot // max(args.steps, 1), "note": "each step times a bounded sample of the step's pairs_per_step_per_gpu pairs"},
            "cpu_baseline": {"value": value, "unit": "stereo frames/s", "cores": cores, "kind": "port",
                             "sample": "%d stereo pairs over %d steps on %d host threads (oracle/frontend_oracle.c; the reference's OpenCV-2.4/stereo-vo path cannot be built here)" % (n_tot, args.steps, cores)},
            "e2e": {"value": value, "unit": "stereo frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--batches-per-step", type=int, default=8, help="batches of --pairs stereo pairs that make one step")
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
    import ctypes
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from srba_stereo_slam_b200 import frontend, synth

    B = args.pairs; BPS = max(1, args.batches_per_step)
    NB = args.steps * BPS                                            # batches in a timed region
    frames = synth.stereo_batch(B, W, H, seed=rank, unique=8)        # [B,2,H,W] uint8: 8 rendered frames x 8 variants, tiled
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=NLEVELS, nfeatures=NFEAT, max_pairs=B, device=local)
    stage = fe.host_staging()
    stage[:2 * B] = frames.reshape(-1, H, W)

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local) if torch.cuda.is_available() else None
        fe.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    # One context replays its batch graph back to back.  (Round 1 alternated two contexts - overlapping tails gave +4-5 % then;
    # with this round's kernels one, two and three contexts measure 108.5 k / 107.3 k / 108.4 k frames/s, tools/multi_ctx_probe.py,
    # profiles/r02k_multi_ctx.json - so the simplest loop is also the fastest.)  One CUDA-event pair on the context's stream spans it.
    clocks = ClockSampler(local); clocks.start()        # spans warm-up, the resident loop and the end-to-end loop
    NCTX = 3
    extra = [frontend.StereoFrontEnd(width=W, height=H, nlevels=NLEVELS, nfeatures=NFEAT, max_pairs=B, device=local) for _ in range(NCTX - 1)]
    rfe = (fe,)
    for f in rfe:
        f.upload(frames)
    for i_ in range(2 * args.warmup * BPS):
        rfe[i_ % len(rfe)].process_device(B)
    for f in rfe:
        f.synchronize()
    kl, kr, dl, dr, m = fe.get_values(0)
    feats = (len(kl), len(kr), len(m))
    barrier()
    l0 = sum(f.kernel_launches() for f in rfe)
    fe.timer_start()
    for i_ in range(NB):
        rfe[i_ % len(rfe)].process_device(B)
    if len(rfe) > 1:
        fe.stream_wait(rfe[1])
    ms = fe.timer_stop()
    rfe[-1].synchronize()
    barrier()
    launches = sum(f.kernel_launches() for f in rfe) - l0
    ms = allmax(ms)
    value = world * B * NB / (ms * 1e-3)

    # ---- parity of the TIMED batch against the oracle (outside the timed region; the oracle is the checker, never the product) ----
    parity = {"n": 0, "ok": 0, "what": []}

    def check(name, cond):
        parity["n"] += 1; parity["ok"] += int(bool(cond))
        if not cond:
            parity["what"].append(name)
    from oracle import oracle as orc
    rng_p = np.random.default_rng(1234 + rank)
    K_PAR = 8 if args.config != "e" else 4
    par_pairs = sorted(int(x) for x in rng_p.choice(B, size=min(K_PAR, B), replace=False))
    oracle_cache = {}

    def oracle_pair(p):
        key = p % max(1, min(B, 64))                                # the batch tiles 64 distinct pairs
        if key not in oracle_cache:
            ol, odl, _ = orc.orb_extract(frames[key, 0], nfeatures=NFEAT, nlevels=NLEVELS, fast_threshold=20)
            orr, odr, _ = orc.orb_extract(frames[key, 1], nfeatures=NFEAT, nlevels=NLEVELS, fast_threshold=20)
            om = orc.match_rowband(ol, odl, orr, odr, max_y_diff=2.0, max_dist=60)
            oracle_cache[key] = (ol, odl, orr, odr, om)
        return oracle_cache[key]
    for p in par_pairs:
        gl, gr, gdl, gdr, gm = rfe[(NB - 1) % len(rfe)].get_values(p)
        ol, odl, orr, odr, om = oracle_pair(p)
        same_kp = len(gl) == len(ol) and len(gr) == len(orr) and all(np.array_equal(gl[f], ol[f]) and np.array_equal(gr[f], orr[f])
                                                                   for f in ("x", "y", "size", "angle", "response", "octave", "class_id"))
        check("pair %d keypoints" % p, same_kp)
        check("pair %d descriptors" % p, same_kp and np.array_equal(gdl, odl) and np.array_equal(gdr, odr))
        check("pair %d L/R matches" % p, len(gm) == len(om) and np.array_equal(gm["queryIdx"], om["queryIdx"]) and np.array_equal(gm["trainIdx"], om["trainIdx"])
              and np.array_equal(gm["distance"], om["distance"]))

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H every batch) -------------------------
    # Three contexts used round-robin (the documented streaming usage of sfe_submit_batch_host / sfe_wait_batch):
    # one batch's PCIe traffic overlaps the other's kernels.  Inputs live in pinned host memory; every batch copies
    # its 2*B images to the device and its keypoints/descriptors/matches/counts back (content-sized, see sfe_last_d2h_bytes).
    for f in extra:
        f.host_staging()[:2 * B] = frames.reshape(-1, H, W)
    fes = (fe,) + tuple(extra)

    def sync_all():
        for f in fes:
            f.synchronize()
        barrier()
    for _ in range(2):
        for f in fes:
            f.submit_staged(B); f.wait_batch()
    sync_all()
    # ---- host-link probe: what the box can move when every rank copies at once and no kernel runs ------------------
    # (a) this rank alone is not measurable under torchrun without serialising the ranks, so: all ranks copy their staging buffer
    # 8 times concurrently, barrier on both sides -> per-rank and aggregate H2D GB/s = the ceiling e2e is held against;
    # (b) the same D2H.
    nprobe = 8
    sync_all(); t0 = time.perf_counter()
    for _ in range(nprobe):
        fe.upload_async(B)
    sync_all(); dt_h2d = allmax(time.perf_counter() - t0)
    h2d_gbs = nprobe * 2 * B * W * H / dt_h2d / 1e9                 # per rank, with all ranks active
    sync_all()
    nmatch = 0; d2h_bytes = 0
    t0 = time.perf_counter()
    for k_ in range(min(NCTX - 1, NB)):
        fes[k_].submit_staged(B)
    for s_ in range(NB):
        if s_ + NCTX - 1 < NB:
            fes[(s_ + NCTX - 1) % NCTX].submit_staged(B)
        f = fes[s_ % NCTX]
        ptrs = f.wait_batch()
        cnt = np.frombuffer((4 * 4 * B * ctypes.c_uint8).from_address(ptrs[3].value), np.int32).reshape(B, 4)
        nmatch += int(cnt[:, 2].sum())                      # the host reads the batch's result
        d2h_bytes += f.last_d2h_bytes()
    sync_all()
    ms_e2e = allmax((time.perf_counter() - t0) * 1e3)
    clk = clocks.stop()
    e2e = world * B * NB / (ms_e2e * 1e-3)
    cap = fe.kp_cap
    h2d = 2 * B * W * H * BPS
    d2h = d2h_bytes // max(args.steps, 1)
    refetch = sum(f.d2h_refetches() for f in fes)
    # e2e parity: the last batch that came back through the pinned result buffers
    last = fes[(NB - 1) % NCTX]
    ptrs = last.wait_batch()
    kp_all = np.frombuffer((ctypes.c_uint8 * (2 * B * cap * 28)).from_address(ptrs[0].value), frontend.KEYPOINT_DTYPE).reshape(2 * B, cap)
    ds_all = np.frombuffer((ctypes.c_uint8 * (2 * B * cap * 32)).from_address(ptrs[1].value), np.uint8).reshape(2 * B, cap, 32)
    mt_all = np.frombuffer((ctypes.c_uint8 * (B * cap * 16)).from_address(ptrs[2].value), frontend.DMATCH_DTYPE).reshape(B, cap)
    cn_all = np.frombuffer((ctypes.c_uint8 * (B * 16)).from_address(ptrs[3].value), frontend.COUNTS_DTYPE)
    for p in par_pairs[:4]:
        ol, odl, orr, odr, om = oracle_pair(p)
        nl, nr, nm_ = int(cn_all[p]["n_left"]), int(cn_all[p]["n_right"]), int(cn_all[p]["n_matches"])
        ok = (nl, nr, nm_) == (len(ol), len(orr), len(om))
        ok = ok and np.array_equal(kp_all[2 * p, :nl], ol) and np.array_equal(kp_all[2 * p + 1, :nr], orr)
        ok = ok and np.array_equal(ds_all[2 * p, :nl], odl) and np.array_equal(ds_all[2 * p + 1, :nr], odr)
        ok = ok and np.array_equal(mt_all[p, :nm_]["trainIdx"], om["trainIdx"]) and np.array_equal(mt_all[p, :nm_]["queryIdx"], om["queryIdx"])
        check("e2e pair %d (pinned result buffers)" % p, ok)

    # ---- full front-end (BASELINE.json configs[2] shape): + BoW query/add + keyframe store + association -----
    full = None
    voc_path = os.path.join(ROOT, "tests", "golden", "voc_flat.npz")
    if os.path.exists(voc_path) and not args.no_full:
        vz = np.load(voc_path)
        voc = {k: (vz[k] if vz[k].ndim else vz[k].item()) for k in vz.files}
        bow = frontend.BowDatabase(fe, voc)
        cap_frames = 4608 if B >= 64 else 36 * B                # >= KITTI-00's 4541 keyframes at the default batch
        kfdb = frontend.KeyframeDatabase(fe, bow, cap_frames)
        nb_full = min(NB, (cap_frames // B) - 4)
        prefill = (cap_frames // B) - nb_full - 1            # untimed batches so the timed ones see a large database
        # resident: two contexts with uploaded frames alternate, so the database work of one batch (latency-bound
        # gathers, on its source context's stream) overlaps the next batch's extraction (issue-bound) on the other stream
        rfe = [fe, fes[1]]
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
        for i_ in range(nb_full):
            f = rfe[i_ % len(rfe)]
            f.process_device(B); kfdb.process_batch(B, src=f)
        for f in rfe:
            f.synchronize()
        ms_full = allmax((time.perf_counter() - t0) * 1e3)
        barrier()
        db1 = kfdb.size()
        # parity of the last timed batch (keyframes db1-B .. db1-1) against the sequential oracle.  All batches hold the same B pairs
        # (64 distinct), so entry e of the database is pair (e % B) % 64.
        ndist = min(B, 64)
        ovec = {}; osub = {}

        def o_vec(k):
            if k not in ovec:
                ol, odl, orr, odr, om = oracle_pair(k)
                w_, wt_ = orc.bow_transform(voc, odl); ovec[k] = orc.bow_vector(voc, w_, wt_); osub[k] = odl[om["queryIdx"]]
            return ovec[k]

        def check_full_batch(res, first_id, tag):
            """res: results of a batch whose frame q has global keyframe id first_id + q; every database entry e holds pair (e % B) % 64."""
            for q in sorted(int(x) for x in rng_p.choice(B, size=min(K_PAR, B), replace=False)):
                g = first_id + q
                qw, qv = o_vec(q % ndist)
                sc_d = {}
                for k in range(ndist):
                    dw, dv = o_vec(k); s_, common = orc.bow_l1_raw(qw, qv, dw, dv); sc_d[k] = (s_, common)
                ent = [(sc_d[(e % B) % ndist][0], e) for e in range(g) if sc_d[(e % B) % ndist][1]]
                ent.sort(); top = [(e, -s_ / 2.0) for s_, e in ent[:4]]
                ok = int(res["query_n"][q]) == len(top) and res["query_ids"][q][:len(top)].tolist() == [t[0] for t in top] \
                    and res["query_scores"][q][:len(top)].tolist() == [t[1] for t in top]
                check("%s frame %d BoW top-4" % (tag, g), ok)
                exp_kf = [g - 1] + [(t[0] if t[0] != g - 1 else -1) for t in top] + [-1] * (4 - len(top))
                okm = res["pair_kf"][q].tolist() == exp_kf
                cur = osub[q % ndist]
                for slot, kf in enumerate(exp_kf):
                    if kf < 0 or not okm:
                        continue
                    o_vec((kf % B) % ndist)
                    obf = orc.match_bf(cur, osub[(kf % B) % ndist])
                    okm = okm and np.array_equal(res["best"][q][slot][:len(cur)], (obf["distance"].astype(np.uint32) << 16) | obf["trainIdx"].astype(np.uint32))
                check("%s frame %d association keys" % (tag, g), okm)
        check_full_batch(kfdb.results(B), db1 - B, "full")
        # end to end: H2D + everything + D2H of the stereo results AND of the query hits / pair list / association keys, EVERY batch
        kfdb.reset()
        nf = len(fes)
        for f in fes:
            f.submit_staged(B); kfdb.process_batch(B, src=f); sl = kfdb.submit_results(); f.wait_batch(); kfdb.wait_results(sl, copy=False)
        nb_e = min(nb_full, (cap_frames - kfdb.size()) // B)
        barrier()
        for f in fes:
            f.synchronize()
        t0 = time.perf_counter()
        slots = {}
        for k_ in range(min(nf - 1, nb_e)):
            fes[k_].submit_staged(B); kfdb.process_batch(B, src=fes[k_]); slots[k_] = kfdb.submit_results()
        hits = 0
        for s_ in range(nb_e):
            if s_ + nf - 1 < nb_e:
                nx = fes[(s_ + nf - 1) % nf]
                nx.submit_staged(B); kfdb.process_batch(B, src=nx); slots[s_ + nf - 1] = kfdb.submit_results()
            fes[s_ % nf].wait_batch()
            hits += int(kfdb.wait_results(slots.pop(s_), copy=False)["query_n"].sum())     # every batch's results are read on the host
        for f in fes:
            f.synchronize()
        barrier()
        ms_full_e2e = allmax((time.perf_counter() - t0) * 1e3)
        full = {"metric": "stereo frames/sec, full front-end (extract + L/R match + BoW query/add + KF match vs previous + 4 hits)",
                "value": world * B * nb_full / (ms_full * 1e-3), "ms_per_batch": ms_full / nb_full,
                "e2e": world * B * nb_e / (ms_full_e2e * 1e-3), "e2e_ms_per_batch": ms_full_e2e / max(nb_e, 1), "query_hits_read_back": hits,
                "db_entries_during_timing": [db0, db1], "batches": nb_full, "e2e_batches": nb_e,
                "database": "per-rank (see full_frontend_sharded for one logical database across the ranks)" if world > 1 else "one",
                "note": "resident: 2 contexts alternate (database work overlaps the next extraction), host wall clock between full synchronisations; "
                        "e2e: 3 contexts; every batch downloads its stereo results and its query hits / pair list / association keys (4 pinned slots)"}

        # ---- configs[2] proper at N > 1: ONE logical database across the ranks, peers read in-kernel over NVLink ------------------------
        if dist is not None:
            kfdb.reset()
            nb_sh = min(nb_full, 16)
            # every rank processes rank 0's frames here: a later shard then finds its best BoW hits among the OLDEST keyframes, which
            # live on other ranks -- loop closures across the shard boundaries, the case the shared database exists for
            fr_t = torch.from_numpy(np.ascontiguousarray(frames)).cuda(); dist.broadcast(fr_t, 0)
            frames_sh = fr_t.cpu().numpy(); del fr_t
            saved_cache = dict(oracle_cache)
            if rank != 0:
                oracle_cache.clear(); ovec.clear(); osub.clear()
            frames_own = frames; frames = frames_sh                 # oracle_pair() reads `frames`
            for f in rfe:
                f.upload(frames_sh)
            starts = kfdb.connect(dist, [nb_sh * B] * world)
            barrier()
            t0 = time.perf_counter()
            for i_ in range(nb_sh):                              # pass 1: extraction + insert, no communication
                f = rfe[i_ % len(rfe)]
                f.process_device(B); kfdb.insert_batch(B, src=f)
            for f in rfe:
                f.synchronize()
            barrier()                                            # every store is complete
            t1 = time.perf_counter()
            for i_ in range(nb_sh):                              # pass 2: query + associate against all OLDER keyframes of all ranks
                kfdb.query_batch(i_ * B, B, src=rfe[i_ % len(rfe)])
            for f in rfe:
                f.synchronize()
            barrier()
            t2 = time.perf_counter()
            ms_sh = allmax((t2 - t0) * 1e3)
            r_sh = kfdb.results(B)
            lo, hi = int(starts[rank]), int(starts[rank + 1])
            check_full_batch(r_sh, hi - B, "sharded rank %d" % rank)
            kfs = r_sh["pair_kf"][r_sh["pair_kf"] >= 0]
            remote = float(((kfs < lo) | (kfs >= hi)).mean()) if len(kfs) else 0.0
            rem = allmax(remote)
            frames = frames_own
            if rank != 0:
                oracle_cache.clear(); oracle_cache.update(saved_cache)
            for f in rfe:
                f.upload(frames)
            full["sharded"] = {"metric": "stereo frames/sec, full front-end, frames sharded over %d GPUs, ONE keyframe database (peer stores read in-kernel over NVLink)" % world,
                               "value": world * B * nb_sh / (ms_sh * 1e-3), "insert_pass_ms": allmax((t1 - t0) * 1e3), "query_pass_ms": allmax((t2 - t1) * 1e3),
                               "frames_total": world * B * nb_sh, "remote_fraction_of_associated_keyframes_last_batch_max_over_ranks": rem,
                               "note": "two passes (insert all, barrier, query all) so that frame g sees exactly the entries < g: results identical to one rank "
                                       "(tests/test_gpu_full_frontend.py::test_one_logical_database_over_ranks_equals_single_rank); checked here against the oracle on every rank"}
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
    # BASELINE.json configs[3] shape: query keyframes x 5 candidates vs a keyframe database sharded over the ranks (peer mode: the
    # candidate blocks are read in-kernel from the owner's memory through CUDA IPC / NVLink; all-gather mode: NCCL replicates the
    # database first) -- run under torchrun at every N so the scaling record carries it.
    from srba_stereo_slam_b200 import loop_closure as lc, sharding
    NQ_M, K_M, NDB_M, NMAX_M = 512 * world, 5, 2048 * world, 500
    lcm = lc.LoopClosureMatcher(fe, NMAX_M, dist)
    s_db, e_db = sharding.shard_range(NDB_M, rank, world); s_q, e_q = sharding.shard_range(NQ_M, rank, world)
    dbc, dbb = lc.synth_keyframe_blocks(e_db - s_db, NMAX_M, seed=100 + rank)
    rng_m = np.random.default_rng(7 + rank)
    src_m = rng_m.integers(0, e_db - s_db, e_q - s_q)
    qc, qb = lc.synth_keyframe_blocks(e_q - s_q, NMAX_M, seed=200 + rank, planted_from=(dbb, dbc, src_m))
    cand_m = rng_m.integers(0, NDB_M, (e_q - s_q, K_M)).astype(np.int32); cand_m[:, 0] = s_db + src_m
    lcm.set_database(dbc, dbb, NDB_M); lcm.upload_queries(qc, qb)
    pr = lcm.build_pairs(cand_m, "peer")

    def timed_lc(mode, gather, reps=10):
        for _ in range(3):
            if gather:
                lcm.replicate_database_allgather()
            lcm.run(mode)
        barrier(); t0_ = time.perf_counter()
        for _ in range(reps):
            if gather:
                lcm.replicate_database_allgather()
            lcm.run(mode)
        fe.synchronize(); barrier()
        return allmax((time.perf_counter() - t0_) * 1e3 / reps)
    ms_m = timed_lc("peer", False)
    best_peer = lcm.results()
    # parity of the matcher run: K pairs against the oracle's brute force
    for i in sorted(int(x) for x in rng_m.choice(e_q - s_q, size=4, replace=False)):
        o = orc.match_bf(qb[i, :qc[i]], dbb[src_m[i], :dbc[src_m[i]]])
        check("loop-closure pair (%d, %d)" % (i, s_db + src_m[i]), np.array_equal(best_peer[i, 0, :qc[i]], (o["distance"].astype(np.uint32) << 16) | o["trainIdx"].astype(np.uint32)))
    lc_block = None
    if dist is not None:
        lcm.replicate_database_allgather(); lcm.build_pairs(cand_m, "allgather")
        ms_ag = timed_lc("allgather", True)
        best_ag = lcm.results()
        check("loop-closure peer == all-gather", np.array_equal(best_peer, best_ag))
        remote_m = float((((cand_m < s_db) | (cand_m >= e_db)) & (cand_m >= 0)).mean())
        lc_block = {"workload": "configs[3] shape: %d query KFs x %d candidates vs a %d-keyframe database of %d descriptors/KF, sharded over %d GPUs" % (NQ_M, K_M, NDB_M, NMAX_M, world),
                    "peer_ms": ms_m, "allgather_ms": ms_ag, "remote_fraction": allmax(remote_m), "query_kf_per_s_peer": NQ_M / (ms_m * 1e-3),
                    "note": "peer: candidate blocks loaded inside the match kernel from the owning GPU (CUDA IPC, NVLink); all-gather: NCCL replicates the "
                            "database, then the same kernel on local memory; identical results; max over ranks, barrier both sides"}
    dist_evals = float((pr["nq"].astype(np.int64) * pr["nt"].astype(np.int64)).sum())
    sm_ghz = (clk.get("sm_mhz") or 1900.0) / 1e3 if isinstance(clk, dict) else 1.9
    POPC_PEAK = 148 * 16 * sm_ghz * 1e9 / 1e12    # measured: 16 POPC/clk/SM (profiles/probes/popc_throughput_probe.cu) x 148 SMs x SM clock under load
    popc_issued = dist_evals * 4 / (ms_m * 1e-3) / 1e12
    matcher = {"bound": "popc pipe", "kernel": "bf_pairs_kernel", "ms_per_launch": ms_m,
               "pairs": int(len(pr)), "distance_evaluations": dist_evals,
               "achieved": popc_issued, "peak": POPC_PEAK, "unit": "T POPC instructions/s", "frac": popc_issued / POPC_PEAK,
               "popc_instructions_per_distance": 4,
               "note": "frac = POPC instructions issued / measured POPC pipe peak (16 per clock per SM).  A 256-bit Hamming distance is 8 popc32 done "
                       "naively; the kernel's carry-save adder tree issues 4 POPC + 16 ALU operations, i.e. %.2f T naive-popc32-equivalents/s" % (2 * popc_issued),
               "workload": "%d query KFs x %d candidates per rank vs a %d-KF database, %d descriptors/KF" % (e_q - s_q, K_M, NDB_M, NMAX_M)}
    lcm.close()

    # ---- per-kernel times, live (CUDA events around each stage, direct launches) --------------------------------
    fe.set_profiling(True)
    acc = {}
    reps = max(3, min(NB, 10))
    for i in range(reps + 1):
        fe.process_device(B); fe.synchronize()
        st = fe.stage_times()
        if i == 0:
            continue
        for k, (t_ms, n) in st.items():
            acc.setdefault(k, [0.0, n]); acc[k][0] += t_ms
    fe.set_profiling(False)
    stages = {k: {"ms": v[0] / reps, "launches": v[1]} for k, v in acc.items() if k not in ("h2d", "d2h") and v[1] > 0}
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
    traffic = None; traffic_src = None
    try:    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this build, scaled to this batch size
        if args.config != "b":
            raise KeyError("traffic was captured for the headline config only")
        tfile = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))[-1]
        with open(os.path.join(ROOT, "profiles", tfile)) as f:
            tj = json.load(f)
        traffic = tj["traffic_bytes_per_launch"][dom] * B / tj["pairs_per_launch"] / max(stages[dom]["launches"], 1)
        traffic_src = "profiles/" + tfile + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one launch; a profiler cannot run inside the bench)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                "stages": stages}

    # ---- CPU baseline (rank 0, N=1 only, bounded sample) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, n, dt = cpu_port_throughput(frames[:4], 1, seconds_budget=12.0)
        cpu = {"value": v, "unit": "stereo frames/s", "cores": 1, "kind": "port",
               "sample": "%d stereo pairs of the same workload in %.1f s, single thread (oracle/frontend_oracle.c)" % (n, dt)}

    pn, pok = parity["n"], parity["ok"]
    if dist is not None:
        t = torch.tensor([pn, pok], device="cuda", dtype=torch.int64); dist.all_reduce(t); pn, pok = int(t[0].item()), int(t[1].item())
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "stereo frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": bench_config(B * BPS),
                "details": {"pairs_per_batch": B, "batches_per_step": BPS, "timed_batches": NB, "timed_region_ms": ms,
                            "keypoints_left_right_matches_pair0": feats,
                            "l2": "per-batch working set %.0f MB (pyramid + blurred arenas of %d images) exceeds the 126 MB L2" % (2 * nimg * sum(level_pixels()) * 1.07 / 1e6, nimg),
                            "parallelism": "frames sharded over %d GPU(s), no data-path collective in the headline path" % world, "rank0_numa_binding": numa},
                "e2e": {"value": e2e, "unit": "stereo frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "matches_read_back": nmatch, "d2h_remainder_fetches": refetch,
                        "d2h_note": "content-sized: %d of %d result slots per image are shipped (%.1f MB per batch instead of %.1f MB)" %
                                    (min(cap, NFEAT + 16), cap, d2h_bytes / max(NB, 1) / 1e6, (2 * B * cap * 60 + B * cap * 16) / 1e6),
                        "h2d_probe": {"GBps_per_gpu_all_ranks_copying": h2d_gbs, "GBps_aggregate": h2d_gbs * world,
                                      "e2e_h2d_GBps_per_gpu": 2 * B * W * H * NB / (ms_e2e * 1e-3) / 1e9,
                                      "e2e_fraction_of_copy_only_ceiling": (2 * B * W * H * NB / (ms_e2e * 1e-3) / 1e9) / h2d_gbs,
                                      "note": "copy-only probe: every rank copies its pinned staging buffer to its GPU 8 times at once, no kernels, barrier both sides"},
                        "timing": "host wall clock between full synchronisations; 3 contexts pipeline PCIe against kernels"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "parity_checked": {"n": pn, "ok": pok, "failed": parity["what"][:8],
                                   "what": "pairs of the TIMED batches vs the CPU oracle, outside the timed regions: keypoints/descriptors/L-R matches (resident + pinned e2e buffers), "
                                           "full front-end BoW top-4 + association keys of the last timed batch, loop-closure matcher pairs"},
                "full_frontend": full, "loop_closure": lc_block, "single_pair_latency": single, "matcher_roofline": matcher}
        print(json.dumps(line))
    fe.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
