#!/usr/bin/env python
"""Benchmark of the wavefront hot path (BASELINE.json metric: alignments/s and
cells-equivalent GCUPS) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--pairs P]
    python bench.py --impl reference ...     # CPU arm: the oracle port on all host cores

A step = one pass of the hot path over one batch of synthetic pairs of the
workload.  `value` = pairs/s with the batch already resident in HBM
(wfacuda_batch_run: pack + align + backtrace kernels, nothing crosses PCIe);
`e2e` = the same metric through wfacuda_align_batch with host buffers, H2D and
D2H inside the timed region.  Under torchrun every rank drives its own GPU
with its own shard of the pair stream (weak scaling, no collective on the data
path); the step time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from wfa_b200 import datagen  # noqa: E402
from wfa_b200 import dist as wdist  # noqa: E402

DEFAULT_WORKLOAD = "cfg2_150bp_e5_global"
WORKLOAD_TEXT = {
    "cfg2_150bp_e5_global": "config 2: 1M synthetic pairs 150 bp, 5% error, global, no heuristic (warp-per-pair path)",
    "cfg3_1kbp_e10_global_adaptive": "config 3: 1M synthetic pairs 1 kbp, 10% error, global, wf-adaptive 10/50",
    "cfg4_10kbp_in_12kbp_e5_semiglobal": "config 4: 100k semi-global alignments, 10 kbp reads vs 12 kbp windows, 5% error",
    "cfg5_100kbp_e15_global_adaptive": "config 5: 10k synthetic pairs 100 kbp, 15% error, global, wf-adaptive 10/50",
}
# bounded CPU samples (about 10-30 s of CPU work on a few dozen cores)
CPU_SAMPLE = {"cfg2_150bp_e5_global": 400_000, "cfg3_1kbp_e10_global_adaptive": 40_000,
              "cfg4_10kbp_in_12kbp_e5_semiglobal": 8, "cfg5_100kbp_e15_global_adaptive": 64}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic(kernel, workload, n_pairs):
    """DRAM bytes per step of the dominant kernel from the committed ncu --set full capture
    (profiles/r1_roofline.json, written by scripts/make_profiles.py); only quoted when this run
    launches the kernel on the same workload and batch size as the capture.  The LANE class is
    several launches per step (the stages of lane_kernel, then lane_finish_kernel): their sum.
    Also returns the executed warp instructions of those launches (INT32-issue evidence)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline.json")))
    except Exception:
        return None, None, None
    size = {"cfg2": ("cfg2_150bp_e5_global", 1_000_000), "cfg3": ("cfg3_1kbp_e10_global_adaptive", 100_000)}
    stem = "lane_" if kernel.startswith("lane") else kernel.split("<")[0]
    traffic = instr = 0
    for key, v in d.get("kernels", {}).items():
        cfg, name = key.split(":", 1)
        if size.get(cfg) == (workload, n_pairs) and stem in name:
            traffic += int(v["dram_bytes"]); instr += int(v["warp_instructions"])
    if not traffic:
        return None, None, None
    return traffic, d.get("source"), instr


def algorithmic_bytes(stats, batch):
    """B = 12*C + ceil((n+m)/4) + 8*R + 64 per pair (2-bit sequences)."""
    return 12 * stats["cells"] + (stats["seq_bases"] + 3) // 4 + 8 * stats["ops"] + 64 * stats["pairs"]


def run_cpu(args, workload, cfgc):
    """CPU arm / baseline: the C restatement of the reference (oracle port; the Go
    reference cannot be built here) on all host cores."""
    import oracle_lib
    cores = os.cpu_count() or 1
    n = args.pairs or CPU_SAMPLE[workload]
    batch = datagen.generate_config(workload, n)
    cfg = oracle_lib.make_config(global_alignment=cfgc["global_alignment"], adaptive=cfgc["adaptive"])
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle_lib.align_batch(cfg, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len,
                               want_ops=True, threads=cores)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n / sec, sec, cores, batch, "%d pairs of the workload per step, %d threads" % (n, cores)


def main():
    # stdout carries exactly one line, the JSON result: libraries that print there (NCCL's version
    # banner, ...) are sent to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wfacuda", choices=["wfacuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(datagen.CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: the config's full size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    workload = args.workload
    cfgc = datagen.CONFIGS[workload]
    rank, world, local_rank = wdist.env()

    if args.impl == "reference":
        if rank != 0:
            return 0
        ws = max(args.warmup, 1) if args.warmup else 0
        args.warmup = min(ws, 1)
        args.steps = min(args.steps, 3)
        v, sec, cores, batch, sample = run_cpu(args, workload, cfgc)
        line = {"impl": "reference", "metric": "alignments_per_sec", "value": v, "unit": "alignments/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "gcups_equiv": batch.cells_equiv() / sec / 1e9,
                "config": {"workload": WORKLOAD_TEXT[workload], "pairs_per_step": len(batch)},
                "cpu_baseline": {"value": v, "unit": "alignments/s", "cores": cores, "kind": "port", "sample": sample,
                                 "note": "C restatement of wfa-go (oracle/); no Go toolchain, reference not buildable"},
                "e2e": {"value": v, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from wfa_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the wfacuda arm has no CPU fallback")
    # ranks of one box share its host cores: give each rank's pipeline its share of them
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world > 1 and "WFACUDA_PIPE_WORKERS" not in os.environ:
        os.environ["WFACUDA_PIPE_WORKERS"] = str(max(2, min(16, (os.cpu_count() or 16) // local_world - 2)))
    torch.cuda.set_device(local_rank)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_pairs = args.pairs or cfgc["pairs"]
    batch = datagen.generate_config(workload, n_pairs, first=wdist.shard_first(rank, n_pairs))   # weak scaling: own shard per rank
    algn = api.New(api.Penalties(4, 6, 2), api.Options(cfgc["global_alignment"]), device=local_rank)
    if cfgc["adaptive"]:
        algn.AdaptiveReduction(api.AdaptiveReductionOption(cfgc["adaptive"][0], cfgc["adaptive"][1], 1))

    # ---- device-resident throughput (value) ---------------------------------
    rb = api.ResidentBatch(algn, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
    for _ in range(args.warmup):
        rb.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ms_align = ms_dev = 0.0
    launches = 0
    for _ in range(args.steps):
        rb.run()
        st = algn.stats()
        ms_align += st["ms_align"]; ms_dev += st["ms_total_device"]; launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    stats = algn.stats()
    results, ops, ops_off = rb.download()
    rb.free()
    ok = int((results["status"] == 0).sum())

    # ---- end to end through the C ABI with host buffers (e2e) ---------------
    # inputs and outputs in page-locked host memory (wfacuda_host_alloc), as a caller that owns
    # its buffers would keep them; H2D of every input and D2H of every result inside the timed region
    e2e_steps = max(args.steps, 20) if n_pairs * 300 <= 400_000_000 and cfgc["pairs"] >= 1_000_000 else max(1, min(args.steps, 5))
    host = [api.pinned_copy(x) for x in (batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)]
    for _ in range(max(args.warmup, 5) if e2e_steps >= 20 else 2):
        algn.align_arrays(*host)    # warm: every pipeline worker has sized its device buffers on a full chunk
    barrier()
    step_ms = []
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        r2, o2, off2 = algn.align_arrays(*host)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    barrier()
    clocks = sampler.stop()        # SM clocks / throttle reasons sampled over both timed regions (resident and e2e)
    # every step is timed on its own (host clock around the blocking call); with >= 10 steps the
    # single slowest one is set aside as a host-scheduling outlier and reported, not hidden
    discarded = None
    kept = list(step_ms)
    if len(kept) >= 10:
        discarded = max(kept)
        kept.remove(discarded)
    wall_e2e = sum(kept) / 1e3
    e2e_counted = len(kept)
    st_e2e = algn.stats()
    assert np.array_equal(r2["score"], results["score"])
    assert np.array_equal(api.ops_in_index_order(r2, o2, off2), api.ops_in_index_order(results, ops, ops_off))
    # the same call on ordinary (pageable) numpy arrays: the library stages them itself
    algn.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
    barrier()
    t2 = time.perf_counter()
    algn.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
    barrier()
    wall_pageable = time.perf_counter() - t2

    # ---- INT32 issue peak of this device, measured ---------------
    int32_peaks = None
    if rank == 0:
        try:
            int32_peaks = algn.measure_int32_peak()
        except Exception:
            int32_peaks = None

    # ---- max over ranks ------------------------------------------------------
    (wall, wall_e2e, ms_align, ms_dev, wall_pageable), (pairs_all, cells_all, ok_all) = wdist.reduce_times_and_totals(
        [wall, wall_e2e, ms_align, ms_dev, wall_pageable], [float(n_pairs), float(batch.cells_equiv()), float(ok)], world, device="cuda")

    if rank == 0:
        sec_step = wall / args.steps
        hbm_peak, peak_src = peaks()
        B = algorithmic_bytes(stats, batch)                      # per launch of the align kernel (this rank)
        k_sec = (ms_align / args.steps) / 1e3
        achieved = B / k_sec / 1e9
        int_ops = 32 * stats["cells"] + 10 * stats["cells"] + 8 * stats["cells"]      # O = 32C + 10V + 8W with V,W ~ C
        kname = max((stats["pairs_lane"], "lane_kernel"), (stats["pairs_warp"], "align_kernel<warp>"), (stats["pairs_cta"], "align_kernel<cta>"),
                    (stats.get("pairs_slim", 0), "slim_kernel"))[1]
        traffic, traffic_src, ncu_instr = ncu_traffic(kname, workload, n_pairs)
        line = {
            "metric": "alignments_per_sec", "value": pairs_all / sec_step, "unit": "alignments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gcups_equiv": cells_all / sec_step / 1e9,
            "config": {"workload": WORKLOAD_TEXT[workload], "pairs_per_gpu_per_step": n_pairs, "penalties": "4/6/2",
                       "global": cfgc["global_alignment"], "adaptive": cfgc["adaptive"],
                       "l2_policy": "inputs+arena larger than L2 (%.0f MB seqs, %.0f MB arena)" % (batch.seq_bytes.nbytes / 1e6, stats["arena_bytes"] / 1e6),
                       "parallelism": "pairs sharded over %d GPU(s), no collective" % world, "pairs_ok": ok_all},
            "e2e": {"value": pairs_all / (wall_e2e / e2e_counted), "unit": "alignments/s",
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]),
                    "gcups_equiv": cells_all / (wall_e2e / e2e_counted) / 1e9, "steps": e2e_counted,
                    "ms_per_step_mean": wall_e2e * 1e3 / e2e_counted, "ms_per_step_median": float(np.median(step_ms)),
                    "ms_per_step_min": min(step_ms), "discarded_slowest_ms": discarded,
                    "host_buffers": "page-locked (wfacuda_host_alloc)", "pageable_value": pairs_all / wall_pageable},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname,
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(B),
                         "kernel_ms": ms_align / args.steps,
                         "cells_per_s": stats["cells"] / k_sec, "int32_ops_per_s_est": int_ops / k_sec},
            # the second bound SURVEY 8(d) names: INT32 issue.  achieved = algorithmic integer ops
            # (O = 32 C + 10 V + 8 W, V and W taken as C) per second of the align phase; peak = one
            # 32-lane integer instruction per scheduler and clock at the SM clock sampled under load
            "roofline_int32": {"bound": "int32-issue", "achieved": int_ops / k_sec / 1e12,
                               "peak": sm_count * 4 * 32 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12, "unit": "Tops/s",
                               "frac": (int_ops / k_sec) / (sm_count * 4 * 32 * (clocks.get("sm_mhz") or 1965.0) * 1e6),
                               # the same peak measured on this device by the library's microbenchmark
                               # (wfacuda_measure_issue_peak): add + xor chains, and add + mad.lo chains
                               "peak_measured": {"add_xor": int32_peaks[0], "add_mad": int32_peaks[1]} if int32_peaks else None,
                               "frac_of_measured": (int_ops / k_sec / 1e12) / max(int32_peaks) if int32_peaks else None,
                               "executed_frac_of_measured": (ncu_instr * 32 / k_sec / 1e12) / max(int32_peaks) if (int32_peaks and ncu_instr) else None,
                               "algorithmic_ops_per_launch": int(int_ops),
                               # what the kernels really issue (committed ncu capture of this workload, all launches of
                               # the class in one step): warp instructions x 32 lanes per second of the align phase
                               "executed_warp_instructions": ncu_instr,
                               "executed_frac": (ncu_instr * 32 / k_sec) / (sm_count * 4 * 32 * (clocks.get("sm_mhz") or 1965.0) * 1e6) if ncu_instr else None,
                               "note": "executed warp instructions and issue-slot utilisation per kernel: profiles/r1_ncu_summary.md"},
            "device_ms_per_step": ms_dev / args.steps,
            "work": {"cells": int(stats["cells"]), "cells_written": int(stats["cells_written"]), "score_steps": int(stats["score_steps"]),
                     "ops": int(stats["ops"]), "retries": int(stats["retries"]), "pairs_lane": int(stats["pairs_lane"]), "pairs_slim": int(stats.get("pairs_slim", 0)), "pairs_warp": int(stats["pairs_warp"]),
                     "pairs_cta": int(stats["pairs_cta"])},
        }
        if not args.no_cpu_baseline and world == 1:
            class A:
                pass
            a = A(); a.pairs = 0; a.warmup = 0; a.steps = 1
            v, sec, cores, cb, sample = run_cpu(a, workload, cfgc)
            line["cpu_baseline"] = {"value": v, "unit": "alignments/s", "cores": cores, "kind": "port", "sample": sample,
                                    "gcups_equiv": cb.cells_equiv() / sec / 1e9}
        emit(line)
    algn.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
