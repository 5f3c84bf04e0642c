#!/usr/bin/env python
"""Turns the ncu reports / launch lists a GPU pass left under gpurun_out/<tag>/ into the tracked
summaries under profiles/ (run here, no GPU needed):  make_profiles.py <tag> [round-prefix]
(prof_cfg2.ncu-rep = the LANE stages + finish kernel of one step, prof_cfg3.ncu-rep = the WARP kernel, launches_cfg2.csv = launch list)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; pre = sys.argv[2] if len(sys.argv) > 2 else "r1"
src = os.path.join(ROOT, "gpurun_out", tag); dst = os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarize(rep, title, fh):
    hdr, units, rows = raw(rep)
    res = {}
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        fh.write("\n### %s — `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % (title, name))
        for i, h in enumerate(hdr):
            if h in KEYS:
                fh.write("| %s | %s | %s |\n" % (h, r[i], units[i]))
            elif 'average_warps_issue_stalled' in h and 'not_issued' not in h:
                try:
                    if float(r[i]) > 0.1:
                        fh.write("| %s | %s | warps per issue-active cycle |\n" % (h.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", ""), r[i]))
                except ValueError:
                    pass
        g = lambda k: (float(r[hdr.index(k)]), units[hdr.index(k)])
        rd, ru = g('dram__bytes_read.sum'); wr, wu = g('dram__bytes_write.sum'); t, tu = g('gpu__time_duration.sum')
        # several launches of one kernel in a step (the LANE stages): summed
        e = res.setdefault(name, {"dram_bytes": 0.0, "duration_ms": 0.0, "warp_instructions": 0.0, "launches": 0})
        e["dram_bytes"] += rd * UNIT[ru] + wr * UNIT[wu]; e["duration_ms"] += t * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[tu]
        e["warp_instructions"] += float(r[hdr.index('smsp__inst_executed.sum')]); e["launches"] += 1
        e["issue_active_pct_last"] = float(r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')])
    return res


os.makedirs(dst, exist_ok=True)
roof = {}
with open(os.path.join(dst, "%s_ncu_summary.md" % pre), "w") as fh:
    fh.write("# ncu summaries (%s), from gpurun_out/%s — `ncu --set full --clock-control none --import-source on`\n" % (pre, tag))
    fh.write("\nCommands: `WFACUDA_NO_PIPELINE=1 ncu --set full --clock-control none --import-source on -k regex:lane_ -s 7 -c 5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (config 2: the four stage launches of `lane_kernel` and `lane_finish_kernel` of one step) and `... -k regex:align_kernel -s 3 -c 1 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 100000 ...` (config 3).  Times under ncu are cold-cache, serialised; bench values come from `bench.py` runs without a profiler.\n")
    for f, title in (("prof_cfg2.ncu-rep", "config 2 (1 M x 150 bp, global) dominant kernel"), ("prof_cfg3.ncu-rep", "config 3 (100 k x 1 kbp, adaptive) dominant kernel")):
        p = os.path.join(src, f)
        if os.path.exists(p):
            roof.update({("cfg2" if "cfg2" in f else "cfg3") + ":" + k: v for k, v in summarize(p, title, fh).items()})
json.dump({"source": "ncu --set full capture, gpurun_out/%s (see %s_ncu_summary.md)" % (tag, pre), "kernels": roof}, open(os.path.join(dst, "%s_roofline.json" % pre), "w"), indent=1)
lp = os.path.join(src, "launches_cfg2.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5 and r[0].isdigit() and "gpu__time_duration" in r[-3]]
    per = {}
    for r in rows:
        per.setdefault(r[4], []).append(float(r[-1]) / 1e6)
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(dst, "%s_launches_cfg2.md" % pre), "w") as fh:
        fh.write("# Launch list, config 2 bench (`ncu --metrics gpu__time_duration.sum --clock-control none`, gpurun_out/%s/launches_cfg2.csv)\n\n" % tag)
        fh.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            fh.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, len(v), sum(v), 100 * sum(v) / tot))
    import shutil
    shutil.copy(lp, os.path.join(dst, "%s_launches_cfg2.csv" % pre))
print("wrote", sorted(os.listdir(dst)))
