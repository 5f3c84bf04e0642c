"""ctypes binding of oracle/libwfaoracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference (oracle/wfa_oracle.h).  It
is loaded only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; the product (wfa_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


class OracleConfig(C.Structure):
    _fields_ = [("mismatch", C.c_uint32), ("gap_open", C.c_uint32), ("gap_ext", C.c_uint32),
                ("global_alignment", C.c_uint8), ("adaptive", C.c_uint8), ("pad_", C.c_uint8 * 2),
                ("min_wf_len", C.c_uint32), ("max_dist_diff", C.c_uint32)]


class OracleResult(C.Structure):
    _fields_ = [("score", C.c_uint32), ("tbegin", C.c_int32), ("tend", C.c_int32),
                ("qbegin", C.c_int32), ("qend", C.c_int32), ("align_len", C.c_uint32),
                ("matches", C.c_uint32), ("gaps", C.c_uint32), ("gap_regions", C.c_uint32),
                ("n_ops", C.c_uint32), ("status", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class OracleCounters(C.Structure):
    _fields_ = [("cells", C.c_uint64), ("visits", C.c_uint64), ("words", C.c_uint64),
                ("ops", C.c_uint64), ("scores", C.c_uint64), ("max_width", C.c_uint64)]


RESULT_DTYPE = np.dtype([("score", "<u4"), ("tbegin", "<i4"), ("tend", "<i4"), ("qbegin", "<i4"),
                         ("qend", "<i4"), ("align_len", "<u4"), ("matches", "<u4"), ("gaps", "<u4"),
                         ("gap_regions", "<u4"), ("n_ops", "<u4"), ("status", "u1"), ("pad_", "u1", 3)])
assert RESULT_DTYPE.itemsize == C.sizeof(OracleResult) == 44


def build(force=False):
    so = os.path.join(ORACLE_DIR, "libwfaoracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("wfa_oracle.c", "wfa_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_new.restype = C.c_void_p
        L.oracle_new.argtypes = [C.POINTER(OracleConfig)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_align.restype = C.c_int
        L.oracle_align.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32,
                                   C.POINTER(OracleResult), C.POINTER(C.POINTER(C.c_uint64)),
                                   C.POINTER(OracleCounters)]
        L.oracle_get_raw.restype = C.c_int
        L.oracle_get_raw.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
        L.oracle_krange.restype = C.c_int
        L.oracle_krange.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_max_score.restype = C.c_uint32
        L.oracle_max_score.argtypes = [C.c_void_p]
        L.oracle_align_batch.restype = C.c_int
        L.oracle_align_batch.argtypes = [C.POINTER(OracleConfig), C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                         C.POINTER(C.c_uint64), C.c_int, C.POINTER(OracleCounters)]
        _LIB = L
    return _LIB


def make_config(mismatch=4, gap_open=6, gap_ext=2, global_alignment=True, adaptive=None):
    """adaptive: None (ad == nil) or (min_wf_len, max_dist_diff)."""
    cfg = OracleConfig()
    cfg.mismatch, cfg.gap_open, cfg.gap_ext = mismatch, gap_open, gap_ext
    cfg.global_alignment = 1 if global_alignment else 0
    cfg.adaptive = 0 if adaptive is None else 1
    if adaptive is not None:
        cfg.min_wf_len, cfg.max_dist_diff = adaptive
    return cfg


def ops_to_cigar(ops):
    return "".join("%d%s" % (int(op) & 0xFFFFFFFF, chr(int(op) >> 32)) for op in ops)


class Oracle:
    """One reference-like Aligner (wfa.go:79-131) backed by the C restatement."""

    def __init__(self, **kw):
        self.cfg = make_config(**kw)
        self.h = lib().oracle_new(C.byref(self.cfg))

    def close(self):
        if self.h:
            lib().oracle_free(self.h)
            self.h = None

    __del__ = close

    def align(self, q, t):
        """-> dict with the AlignmentResult fields, 'ops' (list of int), 'cigar', 'counters'."""
        q, t = bytes(q), bytes(t)
        res, ops, ctr = OracleResult(), C.POINTER(C.c_uint64)(), OracleCounters()
        st = lib().oracle_align(self.h, q, len(q), t, len(t), C.byref(res), C.byref(ops), C.byref(ctr))
        out = {f: getattr(res, f) for f, _ in OracleResult._fields_ if f != "pad_"}
        out["status"] = st
        out["ops"] = [ops[i] for i in range(res.n_ops)] if st == 0 else []
        out["cigar"] = ops_to_cigar(out["ops"])
        out["counters"] = {f: getattr(ctr, f) for f, _ in OracleCounters._fields_}
        return out

    def get_raw(self, comp, s, k):
        raw = C.c_uint32()
        ok = lib().oracle_get_raw(self.h, comp, s, k, C.byref(raw))
        return raw.value if ok else 0

    def krange(self, comp, s):
        lo, hi = C.c_int(), C.c_int()
        if not lib().oracle_krange(self.h, comp, s, C.byref(lo), C.byref(hi)):
            return None
        return lo.value, hi.value

    def max_score(self):
        return lib().oracle_max_score(self.h)


def align_batch(cfg, seq_bytes, q_off, q_len, t_off, t_len, want_ops=True, threads=1):
    """Batch through the oracle. Arrays are numpy (u1, u8, u4, u8, u4).
    -> (results structured array, ops u8 array, ops_off u8 array, counters dict)"""
    n = len(q_len)
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
    t_off = np.ascontiguousarray(t_off, dtype=np.uint64)
    q_len = np.ascontiguousarray(q_len, dtype=np.uint32)
    t_len = np.ascontiguousarray(t_len, dtype=np.uint32)
    results = np.zeros(n, dtype=RESULT_DTYPE)
    ops_off = np.zeros(n, dtype=np.uint64)
    need = C.c_uint64()
    ctr = OracleCounters()
    cap = int((q_len.astype(np.uint64) + t_len.astype(np.uint64)).sum()) + n + 16 if want_ops else 0
    ops = np.zeros(max(cap, 1), dtype=np.uint64)
    rc = lib().oracle_align_batch(C.byref(cfg), n, seq_bytes.ctypes.data, q_off.ctypes.data, q_len.ctypes.data,
                                  t_off.ctypes.data, t_len.ctypes.data, results.ctypes.data,
                                  ops.ctypes.data if want_ops else None, cap, ops_off.ctypes.data,
                                  C.byref(need), threads, C.byref(ctr))
    assert rc == 0, "oracle ops buffer too small"
    counters = {f: getattr(ctr, f) for f, _ in OracleCounters._fields_}
    return results, ops[:need.value] if want_ops else ops[:0], ops_off, counters
