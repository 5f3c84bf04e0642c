"""Deterministic synthetic pair generator (ctypes over csrc/datagen.c).

Workloads follow SURVEY.md section 8(d) / BASELINE.json `configs`:
  config 2: 150 bp, 5 % edits, global, no heuristic
  config 3: 1 kbp, 10 % edits, global, wf-adaptive 10/50
  config 4: 10 kbp read vs 12 kbp window, 5 % edits, semi-global
  config 5: 100 kbp, 15 % edits, global, wf-adaptive 10/50
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwfagen.so")
_SRC = os.path.join(_HERE, "csrc", "datagen.c")
_LIB = None

BASE_SEED = 0x57464100  # ASCII "WFA\0"

# name -> (config number, L, error rate, window, max_start, global, adaptive, full pair count)
CONFIGS = {
    "cfg2_150bp_e5_global": dict(config=2, L=150, err=0.05, window=0, max_start=0,
                                 global_alignment=True, adaptive=None, pairs=1_000_000),
    "cfg3_1kbp_e10_global_adaptive": dict(config=3, L=1000, err=0.10, window=0, max_start=0,
                                          global_alignment=True, adaptive=(10, 50), pairs=1_000_000),
    "cfg4_10kbp_in_12kbp_e5_semiglobal": dict(config=4, L=10_000, err=0.05, window=12_000, max_start=2000,
                                              global_alignment=False, adaptive=None, pairs=100_000),
    "cfg5_100kbp_e15_global_adaptive": dict(config=5, L=100_000, err=0.15, window=0, max_start=0,
                                            global_alignment=True, adaptive=(10, 50), pairs=10_000),
}


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SRC) > os.path.getmtime(_SO):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", _SO, _SRC])
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.wfagen_stride.restype = C.c_uint64
        L.wfagen_stride.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.wfagen_pairs.restype = None
        L.wfagen_pairs.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


class Batch:
    """A batch in the C-ABI layout: one byte pool + per-pair offsets/lengths."""

    def __init__(self, seq_bytes, q_off, q_len, t_off, t_len):
        self.seq_bytes, self.q_off, self.q_len, self.t_off, self.t_len = seq_bytes, q_off, q_len, t_off, t_len

    def __len__(self):
        return len(self.q_len)

    def pair(self, i):
        q = self.seq_bytes[int(self.q_off[i]):int(self.q_off[i]) + int(self.q_len[i])].tobytes()
        t = self.seq_bytes[int(self.t_off[i]):int(self.t_off[i]) + int(self.t_len[i])].tobytes()
        return q, t

    def cells_equiv(self):
        """sum n*m, the GCUPS-equivalent numerator."""
        return int((self.q_len.astype(np.uint64) * self.t_len.astype(np.uint64)).sum())

    def slice(self, a, b):
        return Batch(self.seq_bytes, self.q_off[a:b], self.q_len[a:b], self.t_off[a:b], self.t_len[a:b])

    @staticmethod
    def from_pairs(pairs):
        """pairs: iterable of (query bytes, target bytes)."""
        q_off, q_len, t_off, t_len, chunks, at = [], [], [], [], [], 0
        for q, t in pairs:
            q, t = bytes(q), bytes(t)
            q_off.append(at); q_len.append(len(q)); chunks.append(q); at += len(q)
            t_off.append(at); t_len.append(len(t)); chunks.append(t); at += len(t)
        buf = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy()
        return Batch(buf, np.array(q_off, np.uint64), np.array(q_len, np.uint32),
                     np.array(t_off, np.uint64), np.array(t_len, np.uint32))


def generate(n_pairs, L, err, window=0, max_start=0, config=0, first=0, threads=None, seed=None):
    """Pairs [first, first+n_pairs) of the stream for `config` (see module doc)."""
    lib = _lib()
    nedits = int(round(err * L))
    stride = lib.wfagen_stride(L, nedits, window)
    out = np.zeros(n_pairs * stride + 64, dtype=np.uint8)
    q_off = np.zeros(n_pairs, np.uint64); t_off = np.zeros(n_pairs, np.uint64)
    q_len = np.zeros(n_pairs, np.uint32); t_len = np.zeros(n_pairs, np.uint32)
    base = (BASE_SEED + config) if seed is None else seed
    lib.wfagen_pairs(base, first, n_pairs, L, nedits, window, max_start, out.ctypes.data,
                     q_off.ctypes.data, q_len.ctypes.data, t_off.ctypes.data, t_len.ctypes.data,
                     threads or min(32, os.cpu_count() or 1))
    return Batch(out, q_off, q_len, t_off, t_len)


def generate_config(name, n_pairs=None, first=0, threads=None):
    c = CONFIGS[name]
    return generate(n_pairs if n_pairs is not None else c["pairs"], c["L"], c["err"], c["window"],
                    c["max_start"], c["config"], first, threads)


def read_pair_file(path):
    """The reference CLI's input format (wfa-go/wfa-go.go:157-178): lines come
    in pairs, first byte ('>' query / '<' target) stripped, no upper-casing."""
    pairs = []
    with open(path, "rb") as fh:
        lines = fh.read().split(b"\n")
    i = 0
    while i + 1 < len(lines):
        if lines[i] == b"" and i + 2 >= len(lines):
            break
        pairs.append((lines[i][1:], lines[i + 1][1:]))
        i += 2
    return pairs
