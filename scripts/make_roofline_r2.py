#!/usr/bin/env python
"""profiles/r2_roofline.json from the launch lists of a capture pass (scripts/r2_capture.sh -> gpurun_out/<tag>/launches_*.csv):
per workload, the DRAM bytes and executed warp instructions of the dominant worker class's kernels for ONE step, as
bench.py quotes them in `roofline.traffic` / `roofline_int32.executed_*` -- stamped with the hash of the kernel sources,
so that bench.py stops quoting them the moment a kernel changes.
A launch list holds every launch of one short bench run (3 warm-up steps, 1 timed step, the e2e and api legs): runs of
the batch are told apart by their pack launch, and the full-size steady-state ones averaged.
   make_roofline_r2.py <tag> [note]"""
import collections, csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
SOURCES = ("wfa_kernels.cuh", "wfa_lane.cuh", "wfa_slim.cuh", "wfa_wide.cuh")           # = bench.KERNEL_SOURCES
CLASS = {"cfg2": ("cfg2_150bp_e5_global", 1000000, ("lane_kernel", "lane_finish_kernel", "align_kernel<2, 0>")),
         "cfg3": ("cfg3_1kbp_e10_global_adaptive", 1000000, ("slim_kernel",)),
         "cfg4": ("cfg4_10kbp_in_12kbp_e5_semiglobal", 296, ("wide_kernel", "wide_finish_kernel")),
         "cfg5": ("cfg5_100kbp_e15_global_adaptive", 1250, ("slim_kernel",))}
T, RD, WR, IN = 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum'
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
h = hashlib.sha256()
for f in SOURCES:
    h.update(open(os.path.join(ROOT, "wfa_b200", "csrc", f), "rb").read())
note = sys.argv[2] if len(sys.argv) > 2 else None          # e.g. what changed in the sources since the capture without changing the machine code
out = {"source": "ncu per-launch counters, gpurun_out/%s (scripts/r2_capture.sh; profiles/r2_launches.md)%s" % (tag, "; " + note if note else ""),
       "kernel_source_sha": h.hexdigest()[:16], "workloads": {}}
for short, (wl, pairs, names) in CLASS.items():
    path = os.path.join(src, "launches_%s.csv" % short)
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hd = rows[hi]; kn, mn, mu, mv, idc = (hd.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try: per.setdefault((int(r[idc]), r[kn].split('(')[0].replace('void ', '').replace('wfak::', '')), {})[r[mn]] = float(r[mv].replace(',', '')) * UNIT.get(r[mu], 1.0)
            except ValueError: pass
    runs, cur = [], None                                   # one entry per run of the batch: sums over the class's kernels
    for (i, k), m in per.items():
        if k.startswith("pack_"):
            cur = {"t": 0.0, "bytes": 0.0, "inst": 0.0, "kernels": collections.Counter()}; runs.append(cur)
        elif cur is not None and any(k.startswith(n) for n in names):
            cur["t"] += m[T]; cur["bytes"] += m[RD] + m[WR]; cur["inst"] += m[IN]; cur["kernels"][k] += 1
    runs = [r for r in runs if r["t"] > 0]
    if not runs:
        continue
    # full-size runs (by bytes moved), then the steady state: the ones within 5 % of their median time (the first batch of a
    # ctx runs the LANE class in one stage and tries the narrow SLIM ring first: slower, not what a step of the bench is)
    bmax = max(r["bytes"] for r in runs)
    big = sorted((r for r in runs if r["bytes"] >= 0.8 * bmax), key=lambda r: r["t"])
    tmed = big[len(big) // 2]["t"]
    full = [r for r in big if abs(r["t"] - tmed) <= 0.05 * tmed]
    n = len(full)
    out["workloads"][wl] = {"pairs": pairs, "dram_bytes": int(sum(r["bytes"] for r in full) / n), "warp_instructions": int(sum(r["inst"] for r in full) / n),
                            "class_ms_under_ncu": 1e3 * sum(r["t"] for r in full) / n, "runs_averaged": n,
                            "kernels": dict(full[0]["kernels"])}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_roofline.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
