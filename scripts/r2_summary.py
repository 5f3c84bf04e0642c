#!/usr/bin/env python
"""profiles/r2_summary.md from the round's final passes under gpurun_out/:
   r2_summary.py <final tag (scripts/r2_final.sh)> <multi-entry tag (r2_multi8.sh)> <torchrun tags (r2_multi.sh) ...>"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
fin, me = sys.argv[1], sys.argv[2]
tr = sys.argv[3:]
def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None
d = last_json(os.path.join(G, fin, "bench.json")); ref = last_json(os.path.join(G, fin, "bench_ref.json"))
out = []
w = out.append
w("# Round 2 -- final numbers (one B200 unless said otherwise)\n")
w("`scripts/r2_final.sh %s` on a fresh box: `smoke()`, `pytest -m gpu`, `bench.py --impl reference`, `bench.py` (defaults of the driver: `--gpus 1 --steps 5 --warmup 3`)." % fin)
w("Raw lines: `gpurun_out/%s/bench.json`, `bench_ref.json` (scratch, not tracked); this file is their digest.\n" % fin)
for f in ("smoke.log", "pytest_gpu.log"):
    try:
        w("* `%s`: `%s`" % (f, [l for l in open(os.path.join(G, fin, f)).read().strip().splitlines() if l.strip()][-2 if f.startswith("pytest") else -1].strip()))
    except Exception:
        pass
w("* clocks during the timed regions: %s\n" % json.dumps(d.get("clocks")))
w("## All five configs, one GPU\n")
w("| config | pairs per step | device-resident alignments/s | ms per step | e2e alignments/s (host buffers, copies inside) | dominant kernel | align phase ms | roofline.frac (HBM, algorithmic bytes) | DRAM bytes per step (ncu) / algorithmic | INT32 frac (algorithmic ops / measured issue peak) | CPU port on the box's cores |")
w("|---|---|---|---|---|---|---|---|---|---|---|")
def row(name, c):
    if not isinstance(c, dict) or "value" not in c:
        return
    r = c.get("roofline") or {}; e = c.get("e2e") or {}; i = c.get("roofline_int32") or {}; cpu = c.get("cpu_baseline") or {}
    tr_ = r.get("traffic")
    w("| %s | %s | %.4g | %.3f | %.4g | `%s` | %s | %s | %s | %s | %s |" % (
        name, (c.get("config") or {}).get("pairs_per_gpu_per_step", c.get("pairs_per_step", "")), c["value"], c.get("ms_per_step", 0), e.get("value", 0), r.get("kernel"),
        ("%.3f" % r["kernel_ms"]) if r.get("kernel_ms") else "", ("**%.3f**" % r["frac"]) if r.get("frac") else "",
        ("%.2f GB / %.2f GB = %.2f" % (tr_ / 1e9, r["algorithmic_bytes_per_launch"] / 1e9, tr_ / r["algorithmic_bytes_per_launch"])) if tr_ else "(%s)" % r.get("traffic_source", ""),
        ("%.3f" % i["frac_of_measured"]) if i.get("frac_of_measured") else "", ("%.4g/s on %s threads (%s)" % (cpu["value"], cpu["cores"], cpu["sample"].split(" per")[0])) if cpu else ""))
row("2 (headline)", d)
for k, c in d["configs"].items():
    row(k.split("_")[0].replace("cfg", ""), c)
e = d["e2e"]
w("\nHeadline e2e: %.4g alignments/s = %.2f ms per million pairs (mean of %d steps, min %.2f, max %.2f), %d B in / %d B out per step; pageable host buffers %.4g/s; the reference's call shape (`wfa::Aligner::AlignBatch` on per-pair strings) %.4g/s = %.2f ms per call (%.2fx the C-ABI time)." % (
    e["value"], e["ms_per_step_mean"], e["steps"], e["ms_per_step_min"], e["ms_per_step_max"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], e.get("pageable_value", 0), e.get("api_value", 0), e.get("api_ms_per_call_mean", 0), e.get("api_vs_c_abi", 0)))
if ref:
    w("\nReference arm (`bench.py --impl reference`): %.4g alignments/s on %s threads (%s) -> e2e ratio %.0fx, device-resident ratio %.0fx." % (
        ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"], e["value"] / ref["value"], d["value"] / ref["value"]))
w("\n## Multi-GPU\n")
w("The library's own entry, ONE process (`scripts/r2_multi8.sh`, gpurun_out/%s):\n" % me)
w("| workload | devices | pairs | ms per call (mean / min) | alignments/s | pairs per device |")
w("|---|---|---|---|---|---|")
for wl in ("cfg5_100kbp_e15_global_adaptive", "cfg2_150bp_e5_global"):
    m = last_json(os.path.join(G, me, "me_%s.json" % wl))
    if m:
        m = m["multi_entry"]
        w("| %s | %s | %d | %.2f / %.2f | %.4g | %s |" % (wl, m["entry"].split(", ")[-1], m["pairs"], m["ms_per_call_mean"], m["ms_per_call_min"], m["value"], m["pairs_per_device"]))
if tr:
    w("\nOne process per GPU under torchrun (`scripts/r2_multi.sh`; config 2, 1 M pairs per GPU and step; N > 1 measured earlier in the round: the config-2 kernels changed by less than 1 % since), with the box's measured PCIe ceiling (`scripts/pcie_ceiling.py`: N concurrent streams moving the e2e path's bytes per million pairs, page-locked):\n")
    w("| N | device-resident alignments/s | e2e alignments/s | e2e ms per step (mean) | PCIe ceiling of the box (pairs/s) | e2e / ceiling |")
    w("|---|---|---|---|---|---|")
    ceil = {}
    for t in tr:
        p = os.path.join(G, t, "pcie_ceiling.jsonl")
        if os.path.exists(p):
            for l in open(p):
                try:
                    j = json.loads(l); ceil[j["gpus"]] = j["pairs_per_s_ceiling"]
                except Exception:
                    pass
    rows = {}
    for t in tr + [fin]:
        for f in sorted(os.listdir(os.path.join(G, t))):
            if f.startswith("bench_n") and f.endswith(".json") or (t == fin and f == "bench.json"):
                j = last_json(os.path.join(G, t, f))
                if j and "n_gpus" in j:
                    rows[j["n_gpus"]] = j
    for n in sorted(rows):
        j = rows[n]; c = ceil.get(n)
        w("| %d | %.4g | %.4g | %.2f | %s | %s |" % (n, j["value"], j["e2e"]["value"], j["e2e"]["ms_per_step_mean"], ("%.4g" % c) if c else "", ("%.2f" % (j["e2e"]["value"] / c)) if c else ""))
open(os.path.join(ROOT, "profiles", "r2_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
