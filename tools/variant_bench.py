#!/usr/bin/env python
"""A/B timing of several builds of libsrba_b200.so in one GPU call.

usage: variant_bench.py lib_a.so lib_b.so ...     (paths relative to the repo root; the first one is the reference build)
Each library runs in its own child process on the headline workload (128 pairs of 1241x376, 1000 ORB, 8 levels):
per-stage CUDA-event times (direct launches, median of VB_REPS batches) and a digest of every result array (raw corner
sets per level sorted, keypoints, descriptors, matches).  The parent prints one table and flags any digest that differs
from the first library's.  Variants are built with `make -C srba-stereo-slam_b200/csrc variant NAME=x DEFS=-D...`."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(spec):
    # "path/lib.so@VAR=VALUE" sets an environment knob of the library (e.g. SFE_NO_PDL=1) for this child only
    lib, _, knob = spec.partition("@")
    if knob:
        k, _, v = knob.partition("="); os.environ[k] = v
    sys.path.insert(0, ROOT)
    import numpy as np
    from srba_stereo_slam_b200 import frontend, synth
    frontend.LIB_PATH = os.path.join(ROOT, lib)
    B = int(os.environ.get("VB_PAIRS", "128")); W, H = 1241, 376
    reps = int(os.environ.get("VB_REPS", "12"))
    frames = synth.stereo_batch(B, W, H, seed=0, unique=8)
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=8, nfeatures=1000, max_pairs=B, use_graph=0)
    fe.upload(frames)
    fe.set_profiling(True)
    acc = {}
    for i in range(reps + 2):
        fe.process_device(B); fe.synchronize()
        if i < 2:
            continue
        for k, (t_ms, n) in fe.stage_times().items():
            acc.setdefault(k, []).append(t_ms)
    fe.set_profiling(False)
    med = {k: float(np.median(v)) for k, v in acc.items() if k not in ("h2d", "d2h")}
    # whole-step time with the CUDA graph (what the bench's resident number runs)
    fe2 = frontend.StereoFrontEnd(width=W, height=H, nlevels=8, nfeatures=1000, max_pairs=B, use_graph=1)
    fe2.upload(frames)
    for _ in range(3):
        fe2.process_device(B)
    fe2.synchronize(); fe2.timer_start()
    for _ in range(reps):
        fe2.process_device(B)
    step_ms = fe2.timer_stop() / reps
    h = hashlib.sha256()
    for img in (0, 1, 2 * B - 1):
        for l in range(8):
            x, y, sc = fe.get_raw_corners(img, l)
            h.update(np.sort((sc.astype(np.int64) << 24) | (y.astype(np.int64) << 12) | x).tobytes())
    for p in (0, 1, B // 2, B - 1):
        for a in fe.get_values(p):
            h.update(np.ascontiguousarray(a).tobytes())
    # the CUDA-graph context takes other code paths (forked streams, dependent launches): its results count too
    graph_same = True
    for p in (0, 1, B // 2 - 1, B // 2, B - 2, B - 1):
        for a, b in zip(fe.get_values(p), fe2.get_values(p)):
            graph_same = graph_same and np.array_equal(a, b)
    h.update(b"graph ok" if graph_same else b"GRAPH RESULTS DIFFER FROM THE DIRECT LAUNCHES")
    fe.close(); fe2.close()
    print("VB_RESULT " + json.dumps({"lib": spec, "stages_ms": med, "step_ms": step_ms, "digest": h.hexdigest()[:16] + ("" if graph_same else " GRAPH!=DIRECT")}))


def main():
    libs = sys.argv[1:]
    if os.environ.get("VB_CHILD"):
        return child(libs[0])
    rows = []
    for lib in libs:
        env = dict(os.environ, VB_CHILD="1")
        out = subprocess.run([sys.executable, os.path.abspath(__file__), lib], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("VB_RESULT ")]
        if not line:
            print("FAILED", lib, out.stdout[-2000:], out.stderr[-2000:]); continue
        rows.append(json.loads(line[-1][10:]))
    if not rows:
        sys.exit(1)
    keys = sorted({k for r in rows for k in r["stages_ms"]})
    print("%-44s %9s " % ("library", "step_ms") + " ".join("%13s" % k for k in keys) + "  digest")
    for r in rows:
        flag = "" if r["digest"] == rows[0]["digest"] else "  <-- RESULTS DIFFER"
        print("%-44s %9.4f " % (r["lib"][-44:], r["step_ms"]) + " ".join("%13.4f" % r["stages_ms"].get(k, 0) for k in keys) + "  " + r["digest"] + flag)
    print("VB_JSON " + json.dumps(rows))


if __name__ == "__main__":
    main()
