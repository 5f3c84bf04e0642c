#!/usr/bin/env python
"""Round-2 digests of a capture pass (scripts/r2_capture.sh -> gpurun_out/<tag>/) into profiles/:
   make_profiles_r2.py <tag> [prefix]
 - <prefix>_launches.md: per kernel and workload, launches / time / DRAM bytes / warp instructions / issue / residency
   (from the cheap per-launch metric pass: every launch of one short bench run)
 - <prefix>_ncu_<cfg>.md: key metrics + stall reasons of the `--set full` report of the dominant kernel"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; pre = sys.argv[2] if len(sys.argv) > 2 else "r2"
src = os.path.join(ROOT, "gpurun_out", tag); dst = os.path.join(ROOT, "profiles")
T = 'gpu__time_duration.sum'; RD = 'dram__bytes_read.sum'; WR = 'dram__bytes_write.sum'; IN = 'smsp__inst_executed.sum'
IS = 'smsp__issue_active.avg.pct_of_peak_sustained_active'; WA = 'sm__warps_active.avg.pct_of_peak_sustained_active'
RG = 'launch__registers_per_thread'; TH = 'smsp__thread_inst_executed_per_inst_executed.ratio'
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}

with open(os.path.join(dst, pre + "_launches.md"), "w") as fh:
    fh.write("# Launch lists (%s), gpurun_out/%s\n\n`WFACUDA_NO_PIPELINE=1 ncu --metrics %s --clock-control none --csv python bench.py --workload W --pairs N --steps 1 --warmup 3 --only-headline --no-cpu-baseline` "
             "(scripts/r2_capture.sh).  Every launch of the run (3 warm-up steps + 1 timed + the e2e / api legs), summed per kernel; times under ncu are serialised and cold-cache: "
             "shares and per-launch counters are what to read, bench values come from runs without a profiler.\n" % (pre, tag, ",".join([T, RD, WR, IN, IS, WA, RG, TH])))
    for f in sorted(os.listdir(src)):
        if not (f.startswith("launches_") and f.endswith(".csv")): continue
        rows = list(csv.reader(open(os.path.join(src, f))))
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
        h = rows[hi]; kn, mn, mu, mv, idc = (h.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
        per = collections.defaultdict(dict)
        for r in rows[hi + 1:]:
            if len(r) > mv:
                try: per[(r[idc], r[kn].split('(')[0].replace('void ', ''))][r[mn]] = float(r[mv].replace(',', '')) * UNIT.get(r[mu], 1.0)
                except ValueError: pass
        agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
        for (i, k), m in per.items():
            cnt[k] += 1
            for a, b in m.items(): agg[k][a] += b
        tot = sum(m[T] for m in agg.values())
        fh.write("\n## %s\n\n| kernel | launches | time ms | share | DRAM read GB | DRAM write GB | warp instr (G) | issue active %% | warps active %% | regs | threads/instr |\n|---|---|---|---|---|---|---|---|---|---|---|\n" % f[9:-4])
        for k, m in sorted(agg.items(), key=lambda x: -x[1][T]):
            n = cnt[k]
            if 'int32_peak' in k: continue
            fh.write("| `%s` | %d | %.2f | %.1f %% | %.2f | %.2f | %.3f | %.1f | %.1f | %.0f | %.1f |\n" % (k, n, m[T] * 1e3, 100 * m[T] / tot, m[RD] / 1e9, m[WR] / 1e9, m[IN] / 1e9, m[IS] / n, m[WA] / n, m[RG] / n, m[TH] / n))

KEYS = [T, 'launch__grid_size', 'launch__block_size', RG, 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic', IN, TH, IS, WA,
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', RD, WR, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
for f in sorted(os.listdir(src)):
    if not f.endswith(".ncu-rep"): continue
    out = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
    with open(os.path.join(dst, "%s_ncu_%s.md" % (pre, f[5:-8])), "w") as fh:
        fh.write("# ncu --set full, %s (gpurun_out/%s/%s)\n\n`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 3 -c 1 python bench.py --workload ... --steps 1 --warmup 3` (scripts/r2_capture.sh)\n" % (f[5:-8], tag, f))
        for r in rows[2:]:
            fh.write("\n## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % r[hdr.index("Kernel Name")])
            for i, hh in enumerate(hdr):
                if hh in KEYS: fh.write("| %s | %s | %s |\n" % (hh, r[i], units[i]))
                elif 'average_warps_issue_stalled' in hh and 'not_issued' not in hh:
                    try:
                        if float(r[i]) > 0.1: fh.write("| stall: %s | %s | warps per issue-active cycle |\n" % (hh.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i]))
                    except ValueError: pass
