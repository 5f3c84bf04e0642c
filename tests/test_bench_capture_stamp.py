"""bench.py quotes `roofline.traffic` (DRAM bytes per step from ncu) only from a capture that belongs to the kernels it
runs: profiles/r2_roofline.json is stamped with a hash of the kernel sources and lists the batch size of every captured
workload.  No GPU needed: the mechanism, not the numbers."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_capture_is_quoted_only_for_the_captured_sources_and_batch_size(monkeypatch):
    bench = _bench()
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_roofline.json")))
    wl, k = next(iter(d["workloads"].items()))
    # same sources, same batch size: quoted
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: d["kernel_source_sha"])
    cap, why = bench.ncu_capture(wl, k["pairs"])
    assert why is None and cap["dram_bytes"] == int(k["dram_bytes"]) and cap["warp_instructions"] == int(k["warp_instructions"])
    # another batch size: not quoted, and the reason says so
    cap, why = bench.ncu_capture(wl, k["pairs"] + 1)
    assert cap is None and "batch size" in why
    # a workload that was not captured
    cap, why = bench.ncu_capture("no_such_workload", 1)
    assert cap is None and why
    # the kernels changed since the capture: nothing is replayed
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: "0" * 16)
    cap, why = bench.ncu_capture(wl, k["pairs"])
    assert cap is None and "changed" in why


def test_hash_covers_every_kernel_source():
    bench = _bench()
    have = {f for f in os.listdir(os.path.join(ROOT, "wfa_b200", "csrc")) if f.endswith(".cuh") and f != "wfa_render.cuh"}
    assert have == set(bench.KERNEL_SOURCES), (have, bench.KERNEL_SOURCES)      # (the render kernels are not on the timed path)
    assert len(bench.kernel_source_sha()) == 16
