"""Host-side mirror of the reference's Go API on top of the C ABI (ctypes).

Names, argument meaning and error behaviour follow the reference package
(/root/reference/wfa.go, wfa_cigar.go) so that tests read like the
reference's own usage (README.md:153-216):

    algn = wfa.New(wfa.Penalties(4, 6, 2), wfa.Options(GlobalAlignment=True))
    algn.AdaptiveReduction(wfa.AdaptiveReductionOption(10, 50, 1))
    result = algn.Align(q, t)          # -> AlignmentResult, raises ErrEmptySeq/ErrSeqTooLong
    result.CIGAR(False); result.AlignmentText(q, t, False)
    results, errors = algn.AlignBatch(qs, ts)      # new: many pairs per call
    wfa.RecycleAlignmentResult(result); wfa.RecycleAligner(algn)

This is the ctypes twin of the cgo shim shown in INTEGRATION.md (no Go
toolchain exists in this image).  Every call goes to libwfacuda.so; there is
no CPU fallback: without the built library or without a GPU it raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwfacuda.so")

MaxSeqLen = (1 << 29) - 1                      # wfa.go:190


class WfaError(Exception):
    pass


class _ErrEmptySeq(WfaError):
    pass


class _ErrSeqTooLong(WfaError):
    pass


ErrEmptySeq = _ErrEmptySeq("wfa: invalid empty sequence")                                     # wfa.go:187
ErrSeqTooLong = _ErrSeqTooLong("wfa: sequences longer than %d are not supported" % MaxSeqLen)  # wfa.go:193
ErrResources = WfaError("wfacuda: pair needs more device memory than available")


class Penalties:                                # wfa.go:32-36
    def __init__(self, Mismatch=4, GapOpen=6, GapExt=2):
        self.Mismatch, self.GapOpen, self.GapExt = Mismatch, GapOpen, GapExt


class AdaptiveReductionOption:                  # wfa.go:46-50
    def __init__(self, MinWFLen=10, MaxDistDiff=50, CutoffStep=1):
        self.MinWFLen, self.MaxDistDiff, self.CutoffStep = MinWFLen, MaxDistDiff, CutoffStep


class Options:                                  # wfa.go:64-66
    def __init__(self, GlobalAlignment=True):
        self.GlobalAlignment = GlobalAlignment


DefaultPenalties = Penalties(4, 6, 2)           # wfa.go:39-43
DefaultAdaptiveOption = AdaptiveReductionOption(10, 50, 1)   # wfa.go:56-60
DefaultOptions = Options(True)                  # wfa.go:69-71

OpM, OpD, OpI, OpX, OpH = (ord(c) for c in "MDIXH")   # wfa_cigar.go:60-64
MaskLower32 = 4294967295

FLAG_SEMIGLOBAL_LITERAL = 1
FLAG_FORCE_CTA = 2
FLAG_FORCE_8BIT = 4
FLAG_NO_LANE = 8
FLAG_NO_SLIM = 16
FLAG_NO_WIDE = 32


class _Config(C.Structure):
    _fields_ = [("mismatch", C.c_uint32), ("gap_open", C.c_uint32), ("gap_ext", C.c_uint32),
                ("global_alignment", C.c_uint8), ("adaptive", C.c_uint8), ("reserved_", C.c_uint8 * 2),
                ("min_wf_len", C.c_uint32), ("max_dist_diff", C.c_uint32), ("cutoff_step", C.c_uint32),
                ("flags", C.c_uint32), ("arena_budget_bytes", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("cells", C.c_uint64), ("cells_written", C.c_uint64),
                ("score_steps", C.c_uint64), ("ops", C.c_uint64), ("seq_bases", C.c_uint64),
                ("arena_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("align_launches", C.c_uint32), ("retries", C.c_uint32),
                ("pairs_warp", C.c_uint32), ("pairs_cta", C.c_uint32), ("pairs_8bit", C.c_uint32),
                ("ms_pack", C.c_float), ("ms_align", C.c_float), ("ms_total_device", C.c_float),
                ("pairs_lane", C.c_uint32), ("pairs_slim", C.c_uint32), ("pairs_wide", C.c_uint32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


RESULT_DTYPE = np.dtype([("score", "<u4"), ("tbegin", "<i4"), ("tend", "<i4"), ("qbegin", "<i4"),
                         ("qend", "<i4"), ("align_len", "<u4"), ("matches", "<u4"), ("gaps", "<u4"),
                         ("gap_regions", "<u4"), ("n_ops", "<u4"), ("status", "u1"), ("reserved_", "u1", 3)])

EXPORTS = ["wfacuda_device_count", "wfacuda_create", "wfacuda_destroy", "wfacuda_set_config",
           "wfacuda_align_batch", "wfacuda_last_ops_total", "wfacuda_batch_upload", "wfacuda_batch_run",
           "wfacuda_batch_download", "wfacuda_batch_ops_total", "wfacuda_batch_free",
           "wfacuda_align_batch_multi", "wfacuda_shard_plan", "wfacuda_shard_assign", "wfacuda_get_stats", "wfacuda_last_error",
           "wfacuda_host_alloc", "wfacuda_host_free", "wfacuda_host_register", "wfacuda_host_unregister",
           "wfacuda_batch_render", "wfacuda_last_render_total", "wfacuda_align_components", "wfacuda_measure_issue_peak", "wfacuda_chunk_plan", "wfacuda_wide_plan"]

_LIB = None


def load_library():
    """dlopen libwfacuda.so (never a fallback: a missing library is an error)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("WFACUDA_LIB", LIB_PATH)          # development aid: another build of the same library
    if not os.path.exists(path):
        raise WfaError("libwfacuda.so is not built (run `python -m wfa_b200.build`); there is no CPU fallback")
    L = C.CDLL(path)
    vp, u64, u32p = C.c_void_p, C.c_uint64, C.c_void_p
    L.wfacuda_device_count.restype = C.c_int
    L.wfacuda_create.restype = vp
    L.wfacuda_create.argtypes = [C.c_int, C.POINTER(_Config)]
    L.wfacuda_destroy.argtypes = [vp]
    L.wfacuda_set_config.restype = C.c_int
    L.wfacuda_set_config.argtypes = [vp, C.POINTER(_Config)]
    L.wfacuda_align_batch.restype = C.c_int
    L.wfacuda_align_batch.argtypes = [vp, u64, vp, vp, u32p, vp, u32p, vp, vp, u64, vp]
    L.wfacuda_last_ops_total.restype = u64
    L.wfacuda_last_ops_total.argtypes = [vp]
    L.wfacuda_batch_upload.restype = vp
    L.wfacuda_batch_upload.argtypes = [vp, u64, vp, vp, u32p, vp, u32p]
    L.wfacuda_batch_run.restype = C.c_int
    L.wfacuda_batch_run.argtypes = [vp, vp]
    L.wfacuda_batch_download.restype = C.c_int
    L.wfacuda_batch_download.argtypes = [vp, vp, vp, vp, u64, vp]
    L.wfacuda_batch_ops_total.restype = u64
    L.wfacuda_batch_ops_total.argtypes = [vp]
    L.wfacuda_batch_free.argtypes = [vp, vp]
    L.wfacuda_batch_render.restype = C.c_int
    L.wfacuda_batch_render.argtypes = [vp, vp, C.c_int, vp, u64, vp, vp, vp, u64, vp, vp]
    L.wfacuda_last_render_total.argtypes = [vp, vp, vp]
    L.wfacuda_measure_issue_peak.restype = C.c_int
    L.wfacuda_measure_issue_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.wfacuda_align_components.restype = C.c_int
    L.wfacuda_align_components.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, u64, vp, C.c_uint32, vp, vp, u64, vp]
    L.wfacuda_align_batch_multi.restype = C.c_int
    L.wfacuda_align_batch_multi.argtypes = [C.POINTER(vp), C.c_int, u64, vp, vp, u32p, vp, u32p, vp, vp, u64, vp]
    L.wfacuda_wide_plan.restype = C.c_int
    L.wfacuda_wide_plan.argtypes = [u64, C.c_uint32, u64, vp, vp, vp, vp]
    L.wfacuda_chunk_plan.restype = C.c_int
    L.wfacuda_chunk_plan.argtypes = [u64, u64, C.c_int, vp, C.c_uint32, vp]
    L.wfacuda_shard_plan.restype = C.c_int
    L.wfacuda_shard_plan.argtypes = [C.c_int, u64, u32p, u32p, C.c_int, vp]
    L.wfacuda_shard_assign.restype = C.c_int
    L.wfacuda_shard_assign.argtypes = [C.c_int, u64, u32p, u32p, C.c_int, C.c_int, vp, vp]
    L.wfacuda_get_stats.restype = C.c_int
    L.wfacuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.wfacuda_last_error.restype = C.c_char_p
    L.wfacuda_last_error.argtypes = [vp]
    L.wfacuda_host_alloc.restype = vp
    L.wfacuda_host_alloc.argtypes = [C.c_size_t]
    L.wfacuda_host_free.argtypes = [vp]
    L.wfacuda_host_register.restype = C.c_int
    L.wfacuda_host_register.argtypes = [vp, C.c_size_t]
    L.wfacuda_host_unregister.restype = C.c_int
    L.wfacuda_host_unregister.argtypes = [vp]
    _LIB = L
    return L


class _PinnedBlock:
    """Owner of one wfacuda_host_alloc region; freed when the last numpy view of it dies."""

    def __init__(self, nbytes):
        self._L = load_library()
        self.nbytes = max(int(nbytes), 1)
        self.ptr = self._L.wfacuda_host_alloc(self.nbytes)
        if not self.ptr:
            raise WfaError("wfacuda_host_alloc(%d) failed: %s" % (nbytes, (self._L.wfacuda_last_error(None) or b"").decode()))

    def __del__(self):
        if getattr(self, "ptr", None):
            self._L.wfacuda_host_free(self.ptr)
            self.ptr = None


def pinned_empty(n, dtype):
    """numpy array of n elements in page-locked host memory (wfacuda_host_alloc): the library
    DMAs such arrays directly instead of staging them through its own pinned buffers."""
    dtype = np.dtype(dtype)
    blk = _PinnedBlock(int(n) * dtype.itemsize)
    buf = (C.c_uint8 * blk.nbytes).from_address(blk.ptr)
    buf._owner = blk                                # numpy holds buf, buf holds the block
    return np.frombuffer(buf, dtype=dtype, count=int(n))


def pinned_copy(a):
    a = np.ascontiguousarray(a)
    out = pinned_empty(a.size, a.dtype)
    out[:] = a.reshape(-1)
    return out


def shard_plan(n_shards, q_len, t_len, adaptive):
    """cuts[0..n_shards] of wfacuda_align_batch_multi's static sharding (host logic only)."""
    q_len = np.ascontiguousarray(q_len, np.uint32); t_len = np.ascontiguousarray(t_len, np.uint32)
    cuts = np.zeros(n_shards + 1, np.uint64)
    rc = load_library().wfacuda_shard_plan(n_shards, len(q_len), q_len.ctypes.data, t_len.ctypes.data, int(bool(adaptive)), cuts.ctypes.data)
    if rc != 0:
        raise WfaError("wfacuda_shard_plan failed (%d)" % rc)
    return cuts


def shard_assign(n_shards, q_len, t_len, adaptive, global_alignment=True):
    """(shard of every pair, estimated load per shard) of wfacuda_align_batch_multi's length-binned
    LPT partition (host logic only)."""
    q_len = np.ascontiguousarray(q_len, np.uint32); t_len = np.ascontiguousarray(t_len, np.uint32)
    shard_of = np.zeros(len(q_len), np.uint32); cost = np.zeros(n_shards, np.float64)
    rc = load_library().wfacuda_shard_assign(n_shards, len(q_len), q_len.ctypes.data, t_len.ctypes.data, int(bool(adaptive)),
                                            int(bool(global_alignment)), shard_of.ctypes.data, cost.ctypes.data)
    if rc != 0:
        raise WfaError("wfacuda_shard_assign failed (%d)" % rc)
    return shard_of, cost


def chunk_plan(n_pairs, chunk_pairs, tail_levels=2):
    """Chunk boundaries of the pipelined wfacuda_align_batch (host logic only)."""
    cuts = np.zeros(4096, np.uint64)
    n = C.c_uint32(0)
    rc = load_library().wfacuda_chunk_plan(int(n_pairs), int(chunk_pairs), int(tail_levels), cuts.ctypes.data, len(cuts), C.byref(n))
    if rc != 0:
        raise WfaError("wfacuda_chunk_plan failed (%d)" % rc)
    return cuts[:n.value].copy()


def wide_plan(max_diagonals, seq_entries, smem_per_cta=232448):
    """Geometry of a WIDE launch (host logic only): (cluster CTAs, diagonals per CTA, threads per CTA, shared-memory bytes), or None."""
    c, seg, th, sm = C.c_int(0), C.c_uint32(0), C.c_int(0), C.c_uint64(0)
    rc = load_library().wfacuda_wide_plan(int(max_diagonals), int(seq_entries), int(smem_per_cta), C.byref(c), C.byref(seg), C.byref(th), C.byref(sm))
    return None if rc != 0 else (c.value, seg.value, th.value, sm.value)


def device_count():
    return load_library().wfacuda_device_count()


class AlignmentResult:
    """wfa_cigar.go:29-46 (after process(), :136-214)."""
    __slots__ = ("Ops", "Score", "TBegin", "TEnd", "QBegin", "QEnd", "AlignLen", "Matches", "Gaps", "GapRegions")

    def __init__(self, rec, ops):
        self.Ops = ops                          # numpy uint64 view: op<<32 | n
        self.Score = int(rec["score"])
        self.TBegin, self.TEnd = int(rec["tbegin"]), int(rec["tend"])
        self.QBegin, self.QEnd = int(rec["qbegin"]), int(rec["qend"])
        self.AlignLen, self.Matches = int(rec["align_len"]), int(rec["matches"])
        self.Gaps, self.GapRegions = int(rec["gaps"]), int(rec["gap_regions"])

    def CIGAR(self, onlyAignedRegion=False):    # wfa_cigar.go:236-255
        ops = trimOps(self.Ops) if onlyAignedRegion else self.Ops
        return "".join("%d%s" % (int(op) & MaskLower32, chr(int(op) >> 32)) for op in ops)

    def AlignmentText(self, q, t, onlyAignedRegion=False):      # wfa_cigar.go:259-333
        ops = self.Ops
        if onlyAignedRegion:
            q = q[self.QBegin - 1:self.QEnd]
            t = t[self.TBegin - 1:self.TEnd]
            ops = trimOps(ops)
        Q, A, T = bytearray(), bytearray(), bytearray()
        v = h = 0
        for op in ops:
            n, o = int(op) & MaskLower32, int(op) >> 32
            if o == OpM or o == OpX:
                Q += q[v:v + n]; A += (b"|" if o == OpM else b" ") * n; T += t[h:h + n]; v += n; h += n
            elif o == OpI:
                Q += b"-" * n; A += b" " * n; T += t[h:h + n]; h += n
            elif o == OpD or o == OpH:
                Q += q[v:v + n]; A += b" " * n; T += b"-" * n; v += n
        return bytes(Q), bytes(A), bytes(T)


WAVEFRONT_DTYPE = np.dtype([("score", "<u4"), ("lo", "<i4"), ("hi", "<i4"), ("reserved_", "<u4"), ("first_cell", "<u8")])

# wfa_backtrace_types.go:23-37 and the arrows of wfa_component_plot.go:33-40
wfaTypeBits, wfaTypeMask = 3, 7
wfaUnknown, wfaInsertOpen, wfaInsertExt, wfaDeleteOpen, wfaDeleteExt, wfaMismatch, wfaMatch = range(7)
wfaArrows = ["\u2295", "\u27fc", "\U0001f826", "\u21a7", "\U0001f827", "\u2b02", "\u2b0a"]


class ComponentView:
    """One of Aligner.M / I / D as read back from the GPU: Component (wfa_component.go:37-187)
    restricted to what Plot and tests read -- HasScore, Get, GetRaw, GetAfterDiff, KRange."""

    def __init__(self, is_m):
        self.IsM = is_m
        self.W = {}                             # score -> (lo, hi, raw words of diagonals lo..hi)

    def HasScore(self, s):                      # wfa_component.go:81-86
        return s in self.W

    def KRange(self, s):                        # Lo, Hi of the M wavefront of score s after reduce
        lo, hi, _ = self.W[s]
        return lo, hi

    def GetRaw(self, s, k):                     # wfa_component.go:148-155 (0 = absent)
        wf = self.W.get(s)
        if wf is None or k < wf[0] or k > wf[1]:
            return 0
        return int(wf[2][k - wf[0]])

    def Get(self, s, k):                        # wfa_component.go:142-146: offset, type, ok
        raw = self.GetRaw(s, k)
        return raw >> wfaTypeBits, raw & wfaTypeMask, raw != 0

    def GetAfterDiff(self, s, diff, k):         # wfa_component.go:157-163
        if s < diff:
            return 0, 0, False
        return self.Get(s - diff, k)


class Components:
    """Aligner.M, I, D of one aligned pair (wfacuda_align_components) and the reference's Plot over them."""

    def __init__(self, penalties, rows, cells):
        self.p = penalties
        self.M, self.I, self.D = ComponentView(True), ComponentView(False), ComponentView(False)
        for r in rows:
            lo, hi, a = int(r["lo"]), int(r["hi"]), int(r["first_cell"])
            tri = cells[a:a + 3 * (hi - lo + 1)].reshape(-1, 3)
            for j, comp in enumerate((self.M, self.I, self.D)):
                col = tri[:, j]
                if col.any():                   # a component has the score iff it holds a cell there
                    comp.W[int(r["score"])] = (lo, hi, col.copy())

    def plot_matrix(self, q, t, comp="M", notChangeToMatch=False, maxScore=-1):
        """The matrix Plot fills (wfa_component_plot.go:41-188): mat[v][h] = (score, type) or None."""
        M, I, D, p = self.M, self.I, self.D, self.p
        target = {"M": M, "I": I, "D": D}[comp]
        oe, e, x = p.GapOpen + p.GapExt, p.GapExt, p.Mismatch
        nq, nt = len(q), len(t)
        is_m = M.IsM                            # the reference reads algn.M.IsM, whatever component is plotted (:49)
        mat = [[None] * nt for _ in range(nq)]
        vp = hp = 0                             # declared once for the whole function (:60): values carry over between cells
        for s in sorted(target.W):
            if 0 <= maxScore < s:
                break
            lo, hi, _ = target.W[s]
            for k in range(lo, hi + 1):
                offset, typ, ok = target.Get(s, k)
                if not ok:
                    continue
                h = offset - 1
                v = h - k
                if v < 0 or h < 0 or v >= nq or h >= nt or mat[v][h] is not None:
                    continue
                mat[v][h] = (s, typ)
                if not is_m or q[v] != t[h]:
                    continue
                # where the cell stood before extend (:101-131)
                if typ == wfaInsertExt:
                    offset0 = max(M.GetAfterDiff(s, oe, k - 1)[0], I.GetAfterDiff(s, e, k - 1)[0]) + 1
                elif typ == wfaDeleteExt:
                    offset0 = max(M.GetAfterDiff(s, oe, k + 1)[0], D.GetAfterDiff(s, e, k + 1)[0])
                else:
                    isk = max(M.GetAfterDiff(s, oe, k - 1)[0], I.GetAfterDiff(s, e, k - 1)[0]) + 1
                    dsk = max(M.GetAfterDiff(s, oe, k + 1)[0], D.GetAfterDiff(s, e, k + 1)[0])
                    offset0 = max(isk, dsk, M.GetAfterDiff(s, x, k)[0] + 1)
                h00 = offset0 - 1
                if h == h00:                    # not extended at all
                    continue
                v0, h0 = v, h
                if not notChangeToMatch:
                    mat[v0][h0] = (s, wfaMatch)
                n = 0
                while True:                     # walk the extension backwards (:147-170)
                    h -= 1
                    v -= 1
                    if v < 0 or h < 0:
                        break
                    n += 1
                    if mat[v][h] is not None:
                        continue
                    mat[v][h] = (s, typ) if notChangeToMatch else (s, wfaMatch)
                    vp, hp = v, h
                    if q[v] != t[h] or h == h00:
                        break
                if n == 0:
                    vp, hp = v0, h0
                if not notChangeToMatch:
                    mat[vp][hp] = (s, typ)      # the cell where the run started keeps its own type
        return mat

    def Plot(self, q, t, wtr, comp="M", notChangeToMatch=False, maxScore=-1):
        """(*Aligner).Plot (wfa_component_plot.go:41-209): the tab-separated table, arrows + scores."""
        mat = self.plot_matrix(q, t, comp, notChangeToMatch, maxScore)
        wtr.write("   \t " + "".join("\t%3d" % (h + 1) for h in range(len(t))) + "\n")
        wtr.write("   \t " + "".join("\t%3s" % chr(b) for b in t) + "\n")
        for v, b in enumerate(q):
            cells = "".join("\t  ." if c is None else "\t%s%2d" % (wfaArrows[c[1]], c[0]) for c in mat[v])
            wtr.write("%3d\t%s%s\n" % (v + 1, chr(b), cells))


class RenderedAlignmentResult(AlignmentResult):
    """An AlignmentResult whose CIGAR() / AlignmentText() return what the GPU rendered for the
    `onlyAignedRegion` setting of the batch call (any other setting is formatted on the host)."""
    __slots__ = ("_rendered", "_index", "_trim")

    def __init__(self, rec, ops, rendered, index, trim):
        AlignmentResult.__init__(self, rec, ops)
        self._rendered, self._index, self._trim = rendered, index, trim

    def CIGAR(self, onlyAignedRegion=False):
        if bool(onlyAignedRegion) != self._trim:
            return AlignmentResult.CIGAR(self, onlyAignedRegion)
        return self._rendered.CIGAR(self._index)

    def AlignmentText(self, q, t, onlyAignedRegion=False):
        if bool(onlyAignedRegion) != self._trim:
            return AlignmentResult.AlignmentText(self, q, t, onlyAignedRegion)
        return self._rendered.AlignmentText(self._index)


def ops_in_index_order(results, ops, ops_off):
    """Concatenate every pair's ops slice in pair-index order (the buffer order of pairs is
    unspecified for internally chunked batches)."""
    cnt = np.where(results["status"] == 0, results["n_ops"], 0).astype(np.int64)
    total = int(cnt.sum())
    if total == 0:
        return np.zeros(0, np.uint64)
    excl = np.cumsum(cnt) - cnt
    idx = np.repeat(np.asarray(ops_off).astype(np.int64) - excl, cnt) + np.arange(total, dtype=np.int64)
    return np.asarray(ops)[idx]


def Op(op):                                     # wfa_cigar.go:56-58
    return chr(int(op) >> 32), int(op) & MaskLower32


def trimOps(ops):                               # wfa_cigar.go:217-233
    start = end = -1
    for i in range(len(ops)):
        if int(ops[i]) >> 32 == OpM:
            start = i
            break
    for i in range(len(ops) - 1, -1, -1):
        if int(ops[i]) >> 32 == OpM:
            end = i
            break
    return ops[start:end + 1]


class Aligner:
    """wfa.go:79-268 backed by one wfacuda ctx (one Aligner per thread, wfa.go:73-78)."""

    def __init__(self, p, opt, device=0, flags=0, arena_budget_bytes=0, pinned_outputs=True):
        self.p, self.opt, self.ad = p, opt, None
        self._pinned_out = pinned_outputs
        self._flags, self._budget, self._device = flags, arena_budget_bytes, device
        self._L = load_library()
        cfg = self._config()
        self._ctx = self._L.wfacuda_create(device, C.byref(cfg))
        if not self._ctx:
            raise WfaError((self._L.wfacuda_last_error(None) or b"wfacuda_create failed").decode())

    def _config(self):
        c = _Config()
        c.mismatch, c.gap_open, c.gap_ext = self.p.Mismatch, self.p.GapOpen, self.p.GapExt
        c.global_alignment = 1 if self.opt.GlobalAlignment else 0
        c.adaptive = 0 if self.ad is None else 1
        if self.ad is not None:
            c.min_wf_len, c.max_dist_diff, c.cutoff_step = self.ad.MinWFLen, self.ad.MaxDistDiff, self.ad.CutoffStep
        c.flags, c.arena_budget_bytes = self._flags, self._budget
        return c

    def _err(self):
        return (self._L.wfacuda_last_error(self._ctx) or b"").decode()

    def AdaptiveReduction(self, ad):            # wfa.go:134-140
        if ad.MinWFLen == 0:
            raise WfaError("cutoff step should not be 0")
        self.ad = ad
        cfg = self._config()
        if self._L.wfacuda_set_config(self._ctx, C.byref(cfg)) != 0:
            raise WfaError(self._err())

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.wfacuda_destroy(self._ctx)
            self._ctx = None
        self._bufs = None

    __del__ = close

    # ---- batched entry point (new API) -------------------------------------
    def _out_buffers(self, n, cap):
        """Output buffers are kept across calls (a real caller reuses its result arrays too):
        fresh numpy memory would be page-faulted in inside every call."""
        bufs = getattr(self, "_bufs", None)
        if bufs is None or len(bufs[0]) < n or len(bufs[2]) < cap:
            grow = lambda old, need: max(need, int(1.25 * len(old)) if old is not None else 0)
            sizes = (grow(bufs[0] if bufs else None, n), grow(bufs[1] if bufs else None, n),
                     max(grow(bufs[2] if bufs else None, cap), 1))
            self._bufs = bufs = None
            # large result arrays live in page-locked memory: D2H lands in them directly
            big = self._pinned_out and (sizes[0] * RESULT_DTYPE.itemsize >= (1 << 20) or sizes[2] * 8 >= (1 << 20))
            mk = (lambda k, dt: pinned_empty(k, dt)) if big else (lambda k, dt: np.zeros(k, dt))
            bufs = (mk(sizes[0], RESULT_DTYPE), mk(sizes[1], np.uint64), mk(sizes[2], np.uint64))
            if big:
                for b in bufs:
                    b.view(np.uint8)[:] = 0
            self._bufs = bufs
        return bufs[0][:n], bufs[1][:n], bufs[2]

    def align_arrays(self, seq_bytes, q_off, q_len, t_off, t_len, want_ops=True, copy=False):
        """Raw C-ABI call on numpy arrays -> (results, ops, ops_off).  The returned arrays are
        views of buffers owned by the Aligner, valid until its next call (copy=True detaches)."""
        n = len(q_len)
        seq_bytes = np.ascontiguousarray(seq_bytes, np.uint8)
        q_off = np.ascontiguousarray(q_off, np.uint64); t_off = np.ascontiguousarray(t_off, np.uint64)
        q_len = np.ascontiguousarray(q_len, np.uint32); t_len = np.ascontiguousarray(t_len, np.uint32)
        # ops capacity: the buffers kept from the previous call when there are any (the library
        # reports E_OPS_CAPACITY with the exact need otherwise), else 1/4 op per base -- summing
        # two million-entry length arrays in every call costs more than a tenth of config 2's call
        bufs = getattr(self, "_bufs", None)
        if not want_ops:
            cap = 0
        elif bufs is not None and len(bufs[0]) >= n:
            cap = len(bufs[2])
        else:
            cap = int(q_len.sum(dtype=np.uint64) + t_len.sum(dtype=np.uint64)) // 4 + 16 * n + 64
        while True:
            results, ops_off, ops = self._out_buffers(n, cap)
            rc = self._L.wfacuda_align_batch(self._ctx, n, seq_bytes.ctypes.data, q_off.ctypes.data, q_len.ctypes.data,
                                             t_off.ctypes.data, t_len.ctypes.data, results.ctypes.data,
                                             ops.ctypes.data if want_ops else None, cap, ops_off.ctypes.data)
            if rc == -4:                        # WFACUDA_E_OPS_CAPACITY
                cap = int(self._L.wfacuda_last_ops_total(self._ctx))
                continue
            if rc != 0:
                raise WfaError("wfacuda_align_batch failed (%d): %s" % (rc, self._err()))
            total = int(self._L.wfacuda_last_ops_total(self._ctx)) if want_ops else 0
            if copy:
                return results.copy(), ops[:total].copy(), ops_off.copy()
            return results, ops[:total], ops_off

    def AlignBatch(self, qs, ts):
        """[]*AlignmentResult, []error for many pairs in one call."""
        from .datagen import Batch
        b = Batch.from_pairs(zip(qs, ts))
        results, ops, ops_off = self.align_arrays(b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len, copy=True)
        out, errs = [], []
        for i in range(len(results)):
            st = int(results["status"][i])
            if st == 0:
                a = int(ops_off[i])
                out.append(AlignmentResult(results[i], ops[a:a + int(results["n_ops"][i])]))
                errs.append(None)
            else:
                out.append(None)
                errs.append({1: ErrEmptySeq, 2: ErrSeqTooLong}.get(st, ErrResources))
        return out, errs

    def AlignBatchRendered(self, qs, ts, onlyAignedRegion=False):
        """AlignBatch whose results carry the CIGAR string and the three AlignmentText lines as the
        GPU rendered them (wfacuda_batch_render) instead of formatting them on the host."""
        from .datagen import Batch
        b = Batch.from_pairs(zip(qs, ts))
        rb = ResidentBatch(self, b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)
        try:
            rb.run()
            results, ops, ops_off = rb.download()
            rendered = rb.render(onlyAignedRegion=onlyAignedRegion, text=True)
        finally:
            rb.free()
        out, errs = [], []
        for i in range(len(results)):
            st = int(results["status"][i])
            if st == 0:
                a = int(ops_off[i])
                out.append(RenderedAlignmentResult(results[i], ops[a:a + int(results["n_ops"][i])], rendered, i, bool(onlyAignedRegion)))
                errs.append(None)
            else:
                out.append(None)
                errs.append({1: ErrEmptySeq, 2: ErrSeqTooLong}.get(st, ErrResources))
        return out, errs

    def AlignComponents(self, q, t):
        """Align one pair and read back Aligner.M / I / D (wfa.go:80-86) from the worker's arena
        slot: (AlignmentResult, Components).  What the reference's Plot / Print / GetRaw need."""
        q, t = bytes(q), bytes(t)
        res = np.zeros(1, RESULT_DTYPE)
        ops = np.zeros(len(q) + len(t) + 16, np.uint64)
        rows, cells = np.zeros(1, WAVEFRONT_DTYPE), np.zeros(1, np.uint32)
        n_rows, n_cells = C.c_uint32(0), C.c_uint64(0)
        for _ in range(2):      # first call sizes the buffers, second fills them
            rc = self._L.wfacuda_align_components(self._ctx, q, len(q), t, len(t), res.ctypes.data, ops.ctypes.data, len(ops),
                                                  rows.ctypes.data, len(rows), C.byref(n_rows), cells.ctypes.data, len(cells), C.byref(n_cells))
            if rc != -4:
                break
            rows, cells = np.zeros(max(n_rows.value, 1), WAVEFRONT_DTYPE), np.zeros(max(n_cells.value, 1), np.uint32)
        if rc != 0:
            raise WfaError("wfacuda_align_components failed (%d): %s" % (rc, self._err()))
        st = int(res["status"][0])
        if st != 0:
            raise {1: ErrEmptySeq, 2: ErrSeqTooLong}.get(st, ErrResources)
        comps = Components(self.p, rows[:n_rows.value], cells[:n_cells.value])
        return AlignmentResult(res[0], ops[:int(res["n_ops"][0])].copy()), comps

    def Align(self, q, t):                      # wfa.go:196-198
        res, errs = self.AlignBatch([q], [t])
        if errs[0] is not None:
            raise errs[0]
        return res[0]

    def AlignBatchMulti(self, others, seq_bytes, q_off, q_len, t_off, t_len, want_ops=True):
        """One batch sharded over this Aligner's device and `others` (one ctx per device)."""
        ctxs = [self._ctx] + [o._ctx for o in others]
        arr = (C.c_void_p * len(ctxs))(*ctxs)
        n = len(q_len)
        seq_bytes = np.ascontiguousarray(seq_bytes, np.uint8)
        q_off = np.ascontiguousarray(q_off, np.uint64); t_off = np.ascontiguousarray(t_off, np.uint64)
        q_len = np.ascontiguousarray(q_len, np.uint32); t_len = np.ascontiguousarray(t_len, np.uint32)
        bufs = getattr(self, "_bufs", None)
        if not want_ops:
            cap = 0
        elif bufs is not None and len(bufs[0]) >= n:
            cap = len(bufs[2])
        else:
            cap = int(q_len.sum(dtype=np.uint64) + t_len.sum(dtype=np.uint64)) // 4 + 16 * n + 64
        while True:
            results, ops_off, ops = self._out_buffers(n, cap)        # page-locked, kept across calls
            rc = self._L.wfacuda_align_batch_multi(arr, len(ctxs), n, seq_bytes.ctypes.data, q_off.ctypes.data, q_len.ctypes.data,
                                                   t_off.ctypes.data, t_len.ctypes.data, results.ctypes.data,
                                                   ops.ctypes.data if want_ops else None, cap, ops_off.ctypes.data)
            if rc == -4:
                cap = int(self._L.wfacuda_last_ops_total(self._ctx))
                continue
            if rc != 0:
                raise WfaError("wfacuda_align_batch_multi failed (%d): %s" % (rc, self._err()))
            total = int(self._L.wfacuda_last_ops_total(self._ctx)) if want_ops else 0
            return results, ops[:total], ops_off

    def measure_int32_peak(self):
        """(add/xor only, add + mad.lo) INT32 issue peaks of the device in thread-level Tops/s."""
        a, b = C.c_double(0), C.c_double(0)
        if self._L.wfacuda_measure_issue_peak(self._ctx, C.byref(a), C.byref(b)) != 0:
            raise WfaError(self._err())
        return a.value, b.value

    def stats(self):
        st = Stats()
        self._L.wfacuda_get_stats(self._ctx, C.byref(st))
        return st.as_dict()


def New(p=DefaultPenalties, opt=DefaultOptions, **kw):          # wfa.go:120-131
    return Aligner(p, opt, **kw)


def RecycleAligner(algn):                       # wfa.go:102-116
    if algn is not None:
        algn.close()


def RecycleAlignmentResult(cigar):              # wfa_cigar.go:92-96 (nothing to pool here)
    return None


def RecycleAlignmentText(Q, A, T):              # wfa_cigar.go:346-360
    return None


class RenderedBatch:
    """Strings of a batch as wfacuda_batch_render left them: byte buffers + per-pair offsets."""

    def __init__(self, cigar, cigar_off, cigar_len, text, text_off, text_len):
        self.cigar, self.cigar_off, self.cigar_len = cigar, cigar_off, cigar_len
        self.text, self.text_off, self.text_len = text, text_off, text_len

    def CIGAR(self, i):
        o, n = int(self.cigar_off[i]), int(self.cigar_len[i])
        return self.cigar[o:o + n].tobytes().decode("latin-1")

    def AlignmentText(self, i):
        o, n = int(self.text_off[i]), int(self.text_len[i])
        return tuple(self.text[o + j * n:o + (j + 1) * n].tobytes() for j in range(3))


class ResidentBatch:
    """upload / run / download split (wfacuda_batch_*): keeps a batch in HBM."""

    def __init__(self, algn, seq_bytes, q_off, q_len, t_off, t_len):
        self.algn, self.n = algn, len(q_len)
        self._keep = [np.ascontiguousarray(seq_bytes, np.uint8), np.ascontiguousarray(q_off, np.uint64),
                      np.ascontiguousarray(q_len, np.uint32), np.ascontiguousarray(t_off, np.uint64),
                      np.ascontiguousarray(t_len, np.uint32)]
        s, qo, ql, to, tl = self._keep
        self._b = algn._L.wfacuda_batch_upload(algn._ctx, self.n, s.ctypes.data, qo.ctypes.data, ql.ctypes.data,
                                               to.ctypes.data, tl.ctypes.data)
        if not self._b:
            raise WfaError("wfacuda_batch_upload failed: " + algn._err())

    def run(self):
        rc = self.algn._L.wfacuda_batch_run(self.algn._ctx, self._b)
        if rc != 0:
            raise WfaError("wfacuda_batch_run failed (%d): %s" % (rc, self.algn._err()))

    def download(self, want_ops=True):
        L = self.algn._L
        results = np.zeros(self.n, RESULT_DTYPE)
        ops_off = np.zeros(self.n, np.uint64)
        total = int(L.wfacuda_batch_ops_total(self._b)) if want_ops else 0
        ops = np.empty(max(total, 1), np.uint64)
        rc = L.wfacuda_batch_download(self.algn._ctx, self._b, results.ctypes.data,
                                      ops.ctypes.data if want_ops else None, total, ops_off.ctypes.data)
        if rc != 0:
            raise WfaError("wfacuda_batch_download failed (%d): %s" % (rc, self.algn._err()))
        return results, ops[:total], ops_off

    def render(self, onlyAignedRegion=False, text=True):
        """CIGAR strings and AlignmentText lines of every pair, rendered on the GPU
        (wfacuda_batch_render; wfa_cigar.go:236-333).  Returns a RenderedBatch."""
        L = self.algn._L
        n = self.n
        c_off, c_len = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        t_off, t_len = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        cig, txt = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
        for _ in range(2):      # first call sizes the buffers (WFACUDA_E_OPS_CAPACITY), second fills them
            rc = L.wfacuda_batch_render(self.algn._ctx, self._b, int(bool(onlyAignedRegion)),
                                        cig.ctypes.data, len(cig), c_off.ctypes.data, c_len.ctypes.data,
                                        txt.ctypes.data if text else None, len(txt) if text else 0,
                                        t_off.ctypes.data if text else None, t_len.ctypes.data if text else None)
            if rc != -4:
                break
            a, b = C.c_uint64(0), C.c_uint64(0)
            L.wfacuda_last_render_total(self.algn._ctx, C.byref(a), C.byref(b))
            cig, txt = np.zeros(max(a.value, 1), np.uint8), np.zeros(max(b.value, 1), np.uint8)
        if rc != 0:
            raise WfaError("wfacuda_batch_render failed (%d): %s" % (rc, self.algn._err()))
        return RenderedBatch(cig, c_off, c_len, txt if text else None, t_off, t_len)

    def free(self):
        if self._b:
            self.algn._L.wfacuda_batch_free(self.algn._ctx, self._b)
            self._b = None

    __del__ = free
