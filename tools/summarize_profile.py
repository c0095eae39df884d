#!/usr/bin/env python
"""Turn one round's ncu CSVs into the tracked summaries under profiles/.

usage: summarize_profile.py TAG [bench.json]
  gpurun_out/TAG_launches.csv       `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of bench.py
  gpurun_out/TAG_kernels_full.csv   `ncu --set full --clock-control none --csv --page raw` of one tools/mini_bench.py batch
  bench.json                        the bench line of the same build (live CUDA-event stage times)
writes profiles/TAG_launches.csv, TAG_kernels_full.csv (copies), TAG_launches_summary.md, TAG_traffic.json."""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = {"pyramid_kernel": "pyramid", "fast_kernel": "fast_nms", "select_kernel": "select_harris", "describe_fused_kernel": "orient_brief",
         "describe_kernel": "orient_brief", "blur_kernel": "blur", "rowband_match_kernel": "match_rowband", "rowband_compact_kernel": "match_rowband"}


def table(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi]) and r[0].isdigit()]


def short(name):
    return name.split("(")[0].replace("void ", "").split("<")[0]


def main():
    tag = sys.argv[1]
    bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]) if len(sys.argv) > 2 else None
    src = os.path.join(ROOT, "gpurun_out"); dst = os.path.join(ROOT, "profiles")
    for f in ("_launches.csv", "_kernels_full.csv"):
        shutil.copy(os.path.join(src, tag + f), os.path.join(dst, tag + f))
    # ---- launch list: full 256-image batches only ---------------------------------------------------------------
    hdr, rows = table(os.path.join(src, tag + "_launches.csv"))
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict(); nbatch = 0
    for r in rows:
        k = short(r[ix["Kernel Name"]]); grid = r[ix["Grid Size"]]
        dims = [d.strip() for d in grid.strip("()").split(",")]
        # launches of a full 128-pair batch: image axis 256 (or 128: the pyramid phase runs as two image halves), pair axis 128,
        # the describe kernel's persistent grid (one resident wave; the single-pair launches have ~60 CTAs)
        if k == "pyramid_kernel": full = dims[-1] in ("128", "256")
        elif k.startswith("describe"): full = ("256" in dims) or int(dims[0]) >= 400
        elif k.startswith("rowband"): full = "128" in dims
        else: full = ("256" in dims) and k != "unpack_kernel"
        if k == "fast_kernel" and full:
            nbatch += 1
        if full and k in STAGE:
            e = per.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += float(r[ix["Metric Value"]]) / 1e3
    tot = sum(v[1] for v in per.values()) / max(nbatch, 1)
    live = bench["roofline"]["stages"] if bench else {}
    live_tot = sum(v["ms"] for v in live.values()) * 1e3
    lines = ["# Launch list of `bench.py` (%s), full 128-pair batches only" % tag, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --batches-per-step 1 --no-full --no-cpu-baseline`",
             "(`%s_launches.csv`; the single-pair launches of the latency leg are in the CSV too and are left out here). %d full batches captured." % (tag, nbatch), "",
             "| kernel | launches per batch | us per batch (ncu, cold, serialised) | share of the batch | live CUDA-event stage time (bench, same build) | share |",
             "|---|---|---|---|---|---|"]
    stage_us = collections.OrderedDict()
    for k, (n, us) in per.items():
        stage_us[STAGE[k]] = stage_us.get(STAGE[k], 0.0) + us / nbatch
    seen = set()
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        st = STAGE[k]; lv = live.get(st)
        first = st not in seen; seen.add(st)
        lines.append("| `%s` | %.1f | %.1f | %.1f %% | %s | %s |" % (
            k, n / nbatch, us / nbatch, 100 * us / nbatch / tot,
            ("%.1f us" % (lv["ms"] * 1e3)) if lv and first else ("(with the row above)" if lv else "-"),
            ("%.1f %%" % (100 * lv["ms"] * 1e3 / live_tot)) if lv and first else "-"))
    dom = max(stage_us, key=stage_us.get)
    lines += ["", "Sum: %.1f us per batch under ncu vs %.1f us live (stage events); the dominant kernel is the `%s` stage in both (%.1f %% under ncu%s)." % (
        tot, live_tot, dom, 100 * stage_us[dom] / tot, (", %.1f %% live" % (100 * live[dom]["ms"] * 1e3 / live_tot)) if dom in live else "")]
    open(os.path.join(dst, tag + "_launches_summary.md"), "w").write("\n".join(lines) + "\n")
    # ---- full capture: DRAM traffic and the per-kernel utilisation table -----------------------------------------
    hdr, rows = table(os.path.join(src, tag + "_kernels_full.csv"))
    ix = {h: i for i, h in enumerate(hdr)}
    def num(r, k):
        try:
            return float(r[ix[k]].replace(",", ""))
        except (KeyError, ValueError):
            return 0.0
    traffic = collections.OrderedDict(); us = collections.OrderedDict(); util = []
    for r in rows:
        k = short(r[ix["Kernel Name"]])
        if k not in STAGE:
            continue
        st = STAGE[k]
        traffic[st] = traffic.get(st, 0.0) + num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
        us[st] = us.get(st, 0.0) + num(r, "gpu__time_duration.sum") / 1e3
        util.append({"kernel": k, "grid": r[ix["Grid Size"]], "us": num(r, "gpu__time_duration.sum") / 1e3,
                     "warp_instructions": num(r, "smsp__inst_executed.sum"),
                     "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "lsu_data_pipe_pct": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                     "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "shared_wavefronts": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                     "registers": num(r, "launch__registers_per_thread"),
                     "dram_bytes": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")})
    out = {"source": "profiles/%s_kernels_full.csv (ncu --set full --clock-control none, one direct-launch 128-pair batch of tools/mini_bench.py)" % tag,
           "pairs_per_launch": 128, "traffic_bytes_per_launch": traffic, "ncu_us_per_stage": us, "kernels": util,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per stage (pyramid = its 7 launches together)"}
    if bench:
        out["algorithmic_bytes"] = {k: v["algorithmic_bytes"] for k, v in live.items() if v.get("algorithmic_bytes")}
    json.dump(out, open(os.path.join(dst, tag + "_traffic.json"), "w"), indent=1)
    print("\n".join(lines)); print(json.dumps(traffic))


if __name__ == "__main__":
    main()
