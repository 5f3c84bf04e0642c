#!/usr/bin/env python
"""Benchmark of the wavefront hot path (BASELINE.json metric: alignments/s and
cells-equivalent GCUPS) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--pairs P] [--only-headline]
    python bench.py --impl reference ...     # CPU arm: the oracle port on all host cores
    python bench.py --gpus N --multi-entry   # ONE process drives N GPUs through wfacuda_align_batch_multi

A step = one pass of the hot path over one batch of synthetic pairs of the
workload.  `value` = pairs/s with the batch already resident in HBM
(wfacuda_batch_run: pack + align + backtrace kernels, nothing crosses PCIe);
`e2e` = the same metric through wfacuda_align_batch with host buffers, H2D and
D2H inside the timed region; `e2e.api_value` = the reference's call shape
(AlignBatch on per-pair byte strings -> result objects) through the C++ mirror
of the Go API.  The top-level numbers are the headline workload (config 2
unless --workload says otherwise); `configs` carries value / e2e / roofline /
cpu_baseline of all five BASELINE.json configs (bounded steps for the long
ones).  Under torchrun every rank drives its own GPU with its own shard of the
pair stream (weak scaling, no collective on the data path; config 5 is the
10 000-pair config cut into N shards); step times are the max over ranks.
"""
import argparse
import hashlib
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
    "cfg1_seqs_txt": "config 1: wfa-go CLI on wfa-go/seqs.txt pairs, default gap-affine penalties, global (+ wf-adaptive, the CLI default)",
    "cfg2_150bp_e5_global": "config 2: 1M synthetic pairs 150 bp, 5% error, global, no heuristic (warp-per-pair path)",
    "cfg3_1kbp_e10_global_adaptive": "config 3: 1M synthetic pairs 1 kbp, 10% error, global, wf-adaptive 10/50",
    "cfg4_10kbp_in_12kbp_e5_semiglobal": "config 4: 100k semi-global alignments, 10 kbp reads vs 12 kbp windows, 5% error",
    "cfg5_100kbp_e15_global_adaptive": "config 5: 10k synthetic pairs 100 kbp, 15% error, global, wf-adaptive 10/50",
}
# bounded CPU samples (a few seconds of CPU work each on a few dozen cores)
CPU_SAMPLE = {"cfg2_150bp_e5_global": 400_000, "cfg3_1kbp_e10_global_adaptive": 40_000,
              "cfg4_10kbp_in_12kbp_e5_semiglobal": 4, "cfg5_100kbp_e15_global_adaptive": 32}
# pairs per GPU and step when a workload is a row of `configs` (None: the config's full size;
# config 4 writes 0.27 GB of backtrace arena per pair (8 B per cell; 12 B in round 1), config 5 is one config cut into N shards)
SIDE_PAIRS = {"cfg3_1kbp_e10_global_adaptive": None, "cfg4_10kbp_in_12kbp_e5_semiglobal": 296,
              "cfg5_100kbp_e15_global_adaptive": "strong"}
KERNEL_SOURCES = ("wfa_kernels.cuh", "wfa_lane.cuh", "wfa_slim.cuh", "wfa_wide.cuh")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "wfa_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


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


def ncu_capture(workload, n_pairs):
    """DRAM bytes and executed warp instructions per step of the workload's dominant kernel class,
    from the committed `ncu --set full` capture (profiles/r2_roofline.json, written by
    scripts/make_profiles.py).  Only quoted when the kernels have not changed since the capture
    (hash of the kernel sources) and this run launches the same workload at the same batch size;
    otherwise (None, why)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_roofline.json")))
    except Exception:
        return None, "no committed capture"
    if d.get("kernel_source_sha") != kernel_source_sha():
        return None, "kernel sources changed since the capture (%s)" % d.get("source", "?")
    k = d.get("workloads", {}).get(workload)
    if not k or int(k.get("pairs", -1)) != int(n_pairs):
        return None, "capture is for another batch size"
    return {"dram_bytes": int(k["dram_bytes"]), "warp_instructions": int(k["warp_instructions"]), "kernels": k.get("kernels"),
            "source": d.get("source")}, None


def algorithmic_bytes(stats):
    """B = 12*C + ceil((n+m)/4) + 8*R + 64 per pair (2-bit sequences)."""
    return 12 * stats["cells"] + (stats["seq_bases"] + 3) // 4 + 8 * stats["ops"] + 64 * stats["pairs"]


def run_cpu(workload, n, steps=1, warmup=0):
    """CPU arm / baseline: the C restatement of the reference (oracle port; the Go
    reference cannot be built here) on all host cores."""
    import oracle_lib
    cores = os.cpu_count() or 1
    cfgc = datagen.CONFIGS[workload]
    batch = datagen.generate_config(workload, n)
    cfg = oracle_lib.make_config(global_alignment=cfgc["global_alignment"], adaptive=cfgc["adaptive"])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle_lib.align_batch(cfg, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len,
                               want_ops=True, threads=cores)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": n / sec, "unit": "alignments/s", "cores": cores, "kind": "port",
            "sample": "%d pairs of the workload per step, %d threads" % (n, cores),
            "gcups_equiv": batch.cells_equiv() / sec / 1e9, "ms_per_step": sec * 1e3,
            "note": "C restatement of wfa-go (oracle/); no Go toolchain, reference not buildable"}


def make_aligner(api, workload, device):
    cfgc = datagen.CONFIGS[workload]
    algn = api.New(api.Penalties(4, 6, 2), api.Options(cfgc["global_alignment"]), device=device)
    if cfgc["adaptive"]:
        algn.AdaptiveReduction(api.AdaptiveReductionOption(cfgc["adaptive"][0], cfgc["adaptive"][1], 1))
    return algn


def measure(api, workload, n_pairs, steps, warmup, rank, world, device, barrier, e2e_steps):
    """One workload on this rank's GPU: device-resident steps, then end-to-end steps through the C
    ABI with page-locked host buffers.  Returns this rank's raw numbers (reduced by the caller)."""
    batch = datagen.generate_config(workload, n_pairs, first=wdist.shard_first(rank, n_pairs))   # own shard per rank
    algn = make_aligner(api, workload, device)
    rb = api.ResidentBatch(algn, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
    for _ in range(warmup):
        rb.run()
    barrier()
    t0 = time.perf_counter()
    ms_align = ms_dev = 0.0
    launches = 0
    for _ in range(steps):
        rb.run()
        st = algn.stats()
        ms_align += st["ms_align"]; ms_dev += st["ms_total_device"]; launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    stats = algn.stats()
    results, ops, ops_off = rb.download()
    rb.free()
    ok = int((results["status"] == 0).sum())
    # end to end: inputs and outputs in page-locked host memory (wfacuda_host_alloc), as a caller that
    # owns its buffers keeps them; H2D of every input and D2H of every result inside the timed region
    host = [api.pinned_copy(x) for x in (batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)]
    for _ in range(max(warmup, 5) if e2e_steps >= 20 else 2):
        algn.align_arrays(*host)    # warm: every pipeline worker has sized its device buffers on a full chunk
    barrier()
    step_ms = []
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        r2, o2, off2 = algn.align_arrays(*host)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    barrier()
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
    keep = {"scores": results["score"].copy()}
    algn.close()
    return {"wall": wall, "ms_align": ms_align, "ms_dev": ms_dev, "launches": launches, "stats": stats, "ok": ok,
            "step_ms": step_ms, "st_e2e": st_e2e, "wall_pageable": wall_pageable, "n_pairs": n_pairs,
            "cells_equiv": batch.cells_equiv(), "seq_mb": batch.seq_bytes.nbytes / 1e6, "keep": keep}


def report(workload, m, steps, world, sm_count, clocks, int32_peaks):
    """The JSON fields of one workload from the rank-reduced measurement `m` (see main)."""
    stats = m["stats"]
    cfgc = datagen.CONFIGS[workload]
    sec_step = m["wall"] / steps
    hbm_peak, peak_src = peaks()
    B = algorithmic_bytes(stats)                             # per step of the align kernels (rank 0's shard)
    k_sec = (m["ms_align"] / steps) / 1e3
    achieved = B / k_sec / 1e9
    int_ops = 50 * stats["cells"]                            # O = 32 C + 10 V + 8 W with V, W ~ C
    kname = max((stats["pairs_lane"], "lane_kernel"), (stats["pairs_warp"], "align_kernel<warp>"), (stats["pairs_cta"], "align_kernel<cta>"),
                (stats.get("pairs_slim", 0), "slim_kernel"), (stats.get("pairs_wide", 0), "wide_kernel"))[1]
    cap, why = ncu_capture(workload, m["n_pairs"])
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    issue_peak = sm_count * 4 * 32 * sm_mhz * 1e6
    e2e_mean = sum(m["step_ms"]) / len(m["step_ms"])
    return {
        "value": m["pairs_all"] / sec_step, "unit": "alignments/s", "ms_per_step": sec_step * 1e3, "steps": steps,
        "gcups_equiv": m["cells_all"] / sec_step / 1e9,
        "config": {"workload": WORKLOAD_TEXT[workload], "pairs_per_gpu_per_step": m["n_pairs"], "penalties": "4/6/2",
                   "global": cfgc["global_alignment"], "adaptive": cfgc["adaptive"],
                   "l2_policy": "inputs+arena larger than L2 (%.0f MB seqs, %.0f MB arena)" % (m["seq_mb"], stats["arena_bytes"] / 1e6),
                   "parallelism": "pairs sharded over %d GPU(s), no collective" % world, "pairs_ok": m["ok_all"]},
        "e2e": {"value": m["pairs_all"] / (e2e_mean / 1e3), "unit": "alignments/s",
                "h2d_bytes_per_step": int(m["st_e2e"]["h2d_bytes"]), "d2h_bytes_per_step": int(m["st_e2e"]["d2h_bytes"]),
                "gcups_equiv": m["cells_all"] / (e2e_mean / 1e3) / 1e9, "steps": len(m["step_ms"]),
                "ms_per_step_mean": e2e_mean, "ms_per_step_median": float(np.median(m["step_ms"])),
                "ms_per_step_min": min(m["step_ms"]), "ms_per_step_max": max(m["step_ms"]),
                "host_buffers": "page-locked (wfacuda_host_alloc)", "pageable_value": m["pairs_all"] / m["wall_pageable"]},
        "gpu_launches": int(m["launches"]),
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": cap["dram_bytes"] if cap else None, "traffic_source": cap["source"] if cap else why,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": int(B), "kernel_ms": m["ms_align"] / steps,
                     "cells_per_s": stats["cells"] / k_sec},
        # the second bound SURVEY 8(d) names: INT32 issue.  achieved = algorithmic integer ops (50 per
        # wavefront cell) per second of the align phase; peak = one 32-lane integer instruction per
        # scheduler and clock at the SM clock sampled under load, and the same measured on this device
        "roofline_int32": {"bound": "int32-issue", "achieved": int_ops / k_sec / 1e12, "peak": issue_peak / 1e12, "unit": "Tops/s",
                           "frac": (int_ops / k_sec) / issue_peak,
                           "peak_measured": {"add_xor": int32_peaks[0], "add_mad": int32_peaks[1]} if int32_peaks else None,
                           "frac_of_measured": (int_ops / k_sec / 1e12) / max(int32_peaks) if int32_peaks else None,
                           "algorithmic_ops_per_launch": int(int_ops),
                           "executed_warp_instructions": cap["warp_instructions"] if cap else None,
                           "executed_frac": (cap["warp_instructions"] * 32 / k_sec) / issue_peak if cap else None},
        "device_ms_per_step": m["ms_dev"] / steps,
        "work": {k: int(stats.get(k, 0)) for k in ("cells", "cells_written", "score_steps", "ops", "retries", "pairs_lane", "pairs_slim", "pairs_wide", "pairs_warp", "pairs_cta")},
    }


def config1(api, device):
    """Config 1: the reference CLI's own sample pairs (wfa-go/seqs.txt, committed with their README
    output in tests/golden/readme_vectors.json) with the CLI's defaults: global + wf-adaptive 10/50."""
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "readme_vectors.json")))
    pairs = [(p["q"].encode(), p["t"].encode()) for p in G["seqs_txt"]]
    algn = api.New(api.Penalties(4, 6, 2), api.Options(True), device=device)
    algn.AdaptiveReduction(api.AdaptiveReductionOption(10, 50, 1))
    qs, ts = [p[0] for p in pairs], [p[1] for p in pairs]
    for _ in range(3):
        res, errs = algn.AlignBatch(qs, ts)
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        res, errs = algn.AlignBatch(qs, ts)
    sec = (time.perf_counter() - t0) / K
    ok = res[0].CIGAR(False) == "1X1I14M1D39M1D31M1D12M" and res[0].Score == 36          # README.md:245-254
    algn.close()
    return {"value": len(pairs) / sec, "unit": "alignments/s", "ms_per_step": sec * 1e3, "pairs_per_step": len(pairs),
            "config": {"workload": WORKLOAD_TEXT["cfg1_seqs_txt"]}, "matches_readme_output": bool(ok),
            "e2e": {"value": len(pairs) / sec, "unit": "alignments/s", "note": "two pairs: one AlignBatch call = one launch chain, latency-bound"}}


def multi_entry(api, workload, n_pairs, n_dev, steps, check_scores=None):
    """ONE process, one ctx per device, one call: wfacuda_align_batch_multi (LPT shards, the chunked
    pipeline on every device).  End to end with page-locked host buffers."""
    batch = datagen.generate_config(workload, n_pairs)
    algns = [make_aligner(api, workload, d) for d in range(n_dev)]
    host = [api.pinned_copy(x) for x in (batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)]
    try:
        for _ in range(2):
            r, o, off = algns[0].AlignBatchMulti(algns[1:], *host)
        ms = []
        for _ in range(steps):
            t0 = time.perf_counter()
            r, o, off = algns[0].AlignBatchMulti(algns[1:], *host)
            ms.append((time.perf_counter() - t0) * 1e3)
        per_dev = [int(a.stats()["pairs"]) for a in algns]
        same = None if check_scores is None else bool(np.array_equal(r["score"][:len(check_scores)], check_scores))
        ok = int((r["status"] == 0).sum())
    finally:
        for a in algns:
            a.close()
    mean = sum(ms) / len(ms)
    return {"entry": "wfacuda_align_batch_multi, one process, %d devices" % n_dev, "workload": WORKLOAD_TEXT[workload], "pairs": n_pairs,
            "value": n_pairs / (mean / 1e3), "unit": "alignments/s", "ms_per_call_mean": mean, "ms_per_call_min": min(ms), "steps": steps,
            "gcups_equiv": batch.cells_equiv() / (mean / 1e3) / 1e9, "pairs_per_device": per_dev,
            "pairs_ok": ok, "scores_equal_single_device_run": same}


def api_level(workload, n_pairs, device, steps=10, warmup=3):
    """The reference's call shape through the C++ mirror of the Go API (wfa_b200/host/bench_api)."""
    exe = os.path.join(ROOT, "wfa_b200", "host", "bench_api")
    c = datagen.CONFIGS[workload]
    if not os.path.exists(exe) or c["window"]:
        return None
    try:
        out = subprocess.run([exe, str(c["config"]), str(c["L"]), str(int(round(c["err"] * c["L"]))), str(n_pairs),
                              "1" if c["global_alignment"] else "0", "1" if c["adaptive"] else "0", str(steps), str(warmup), str(device)],
                             capture_output=True, text=True, timeout=600)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"error": str(e)[:200]}


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
    ap.add_argument("--only-headline", action="store_true", help="skip the `configs` rows of the other workloads")
    ap.add_argument("--multi-entry", action="store_true", help="one process, --gpus devices, through wfacuda_align_batch_multi")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    workload = args.workload
    cfgc = datagen.CONFIGS[workload]
    rank, world, local_rank = wdist.env()

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warmup = min(args.steps, 3), min(max(args.warmup, 0), 1)
        cb = run_cpu(workload, args.pairs or CPU_SAMPLE[workload], steps, warmup)
        line = {"impl": "reference", "metric": "alignments_per_sec", "value": cb["value"], "unit": "alignments/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "gcups_equiv": cb["gcups_equiv"],
                "config": {"workload": WORKLOAD_TEXT[workload], "pairs_per_step": args.pairs or CPU_SAMPLE[workload]},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from wfa_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the wfacuda arm has no CPU fallback")
    # ranks of one box share its host cores: the pipeline workers of every rank sleep in their waits
    # instead of spinning, so that each rank can keep enough chunks in flight
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world > 1:
        os.environ.setdefault("WFACUDA_PIPE_WORKERS", str(max(4, min(12, 2 * (os.cpu_count() or 16) // local_world))))
        os.environ.setdefault("WFACUDA_BLOCKING_SYNC", "1")
    torch.cuda.set_device(local_rank)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduced(workload_, n_pairs_, steps_, e2e_steps_):
        m = measure(api, workload_, n_pairs_, steps_, args.warmup, rank, world, local_rank, barrier, e2e_steps_)
        e2e_mean = sum(m["step_ms"]) / len(m["step_ms"])
        (wall, e2e_mean_max, ms_align, ms_dev, wall_pageable), (pairs_all, cells_all, ok_all) = wdist.reduce_times_and_totals(
            [m["wall"], e2e_mean, m["ms_align"], m["ms_dev"], m["wall_pageable"]],
            [float(m["n_pairs"]), float(m["cells_equiv"]), float(m["ok"])], world, device="cuda")
        # every step of every rank stays in the mean: the slowest rank's mean is the job's
        scale = e2e_mean_max / e2e_mean if e2e_mean > 0 else 1.0
        m.update(wall=wall, ms_align=ms_align, ms_dev=ms_dev, wall_pageable=wall_pageable, pairs_all=pairs_all, cells_all=cells_all,
                 ok_all=ok_all, step_ms=[x * scale for x in m["step_ms"]])
        return m

    # ---- single process, several devices, through the library's own multi-GPU entry -------------
    if args.multi_entry and world == 1 and args.gpus > 1:
        n_pairs = args.pairs or cfgc["pairs"]
        sampler = ClockSampler(0); sampler.start()
        me = multi_entry(api, workload, n_pairs, args.gpus, max(args.steps, 3))
        clocks = sampler.stop()
        line = {"metric": "alignments_per_sec", "value": me["value"], "unit": "alignments/s", "n_gpus": args.gpus, "steps": me["steps"],
                "warmup": 2, "ms_per_step": me["ms_per_call_mean"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "gcups_equiv": me["gcups_equiv"],
                "config": {"workload": WORKLOAD_TEXT[workload], "pairs_per_step": n_pairs, "parallelism": me["entry"]},
                "e2e": {"value": me["value"], "unit": "alignments/s", "note": "the call is end to end: host buffers in, host buffers out"},
                "multi_entry": me, "clocks": clocks}
        emit(line)
        return 0

    # ---- headline workload ---------------------------------------------------------------------
    n_pairs = args.pairs or cfgc["pairs"]
    if workload == "cfg5_100kbp_e15_global_adaptive" and not args.pairs:
        n_pairs = cfgc["pairs"] // world
    e2e_steps = max(args.steps, 20) if n_pairs * 300 <= 400_000_000 and cfgc["pairs"] >= 1_000_000 and cfgc["L"] <= 200 else max(2, min(args.steps, 5))
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = reduced(workload, n_pairs, args.steps, e2e_steps)
    clocks = sampler.stop()        # SM clocks / throttle reasons sampled over both timed regions (resident and e2e)

    int32_peaks = None
    if rank == 0:
        try:
            a = make_aligner(api, workload, local_rank)
            int32_peaks = a.measure_int32_peak()
            a.close()
        except Exception:
            int32_peaks = None

    # ---- the other configs of BASELINE.json (bounded steps), same ranks, same shards -------------
    side = {}
    if not args.only_headline and not args.pairs:
        for wl in ("cfg3_1kbp_e10_global_adaptive", "cfg4_10kbp_in_12kbp_e5_semiglobal", "cfg5_100kbp_e15_global_adaptive", "cfg2_150bp_e5_global"):
            if wl == workload:
                continue
            sp = SIDE_PAIRS.get(wl)
            np_ = datagen.CONFIGS[wl]["pairs"] if sp is None else (datagen.CONFIGS[wl]["pairs"] // world if sp == "strong" else sp)
            steps_ = 2 if wl != "cfg2_150bp_e5_global" else 5
            side[wl] = (reduced(wl, np_, steps_, 3 if wl != "cfg2_150bp_e5_global" else 10), steps_)

    line = None
    if rank == 0:
        line = {"metric": "alignments_per_sec", "n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic"}
        line.update(report(workload, head, args.steps, world, sm_count, clocks, int32_peaks))
        line["clocks"] = clocks
        line["kernel_source_sha"] = kernel_source_sha()
        configs = {}
        try:
            configs["cfg1_seqs_txt"] = config1(api, local_rank)
        except Exception as e:
            configs["cfg1_seqs_txt"] = {"error": str(e)[:200]}
        configs[workload] = "headline (top-level fields of this line)"
        for wl, (m, steps_) in side.items():
            r = report(wl, m, steps_, world, sm_count, clocks, int32_peaks)
            if SIDE_PAIRS.get(wl) == "strong":
                r["scaling"] = "strong (the 10 000 pairs of the config cut into %d shards)" % world
            elif isinstance(SIDE_PAIRS.get(wl), int):
                r["bounded"] = "%d of the config's %d pairs per step (0.27 GB of backtrace arena per pair)" % (SIDE_PAIRS[wl], datagen.CONFIGS[wl]["pairs"])
            configs[wl] = r
        # CPU port beside every GPU number, on this box's cores (bounded samples), at every N
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = run_cpu(workload, CPU_SAMPLE[workload])
            for wl in side:
                configs[wl]["cpu_baseline"] = run_cpu(wl, CPU_SAMPLE[wl])
        # the reference's call shape (per-pair byte strings -> result objects) on the headline workload
        al = api_level(workload, n_pairs, local_rank)
        if al and "api_value" in al:
            line["e2e"].update(api_value=al["api_value"], api_ms_per_call_mean=al["ms_per_call_mean"], api_ms_per_call_min=al["ms_per_call_min"],
                               api_call=al["call"], api_vs_c_abi=al["ms_per_call_mean"] / line["e2e"]["ms_per_step_mean"])
        elif al:
            line["e2e"]["api_error"] = al.get("error")
        line["configs"] = configs
    if world > 1:
        barrier()
    # ---- the library's own multi-GPU entry from ONE process (rank 0 drives all devices, the other ranks wait)
    if world > 1 and not args.only_headline and not args.pairs:
        # the waiting ranks wait on the CPU (a gloo group): an NCCL barrier would keep a spinning kernel on every other
        # rank's GPU -- the devices rank 0 is about to drive (measured at N = 2: 530 ms per config-5 call with the NCCL
        # barrier, 250-290 ms for the same call from a process of its own)
        cpu_group = dist.new_group(backend="gloo")
        torch.cuda.empty_cache()
        if rank == 0:
            try:
                c5 = "cfg5_100kbp_e15_global_adaptive"
                line["multi_entry"] = {
                    "cfg5": multi_entry(api, c5, datagen.CONFIGS[c5]["pairs"], world, 2, side[c5][0]["keep"]["scores"] if c5 in side else None),
                    "cfg2": multi_entry(api, "cfg2_150bp_e5_global", 1_000_000 * world, world, 5)}
            except Exception as e:
                line["multi_entry"] = {"error": str(e)[:300]}
        dist.barrier(group=cpu_group)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
