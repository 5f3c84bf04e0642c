"""Second, independent restatement of the reference's wavefront path (pure Python).

TEST INFRASTRUCTURE ONLY -- never imported by wfa_b200/.  It exists because
the reference cannot be executed here (no Go toolchain): two restatements
written separately from the Go source (this one and oracle/wfa_oracle.c) must
agree on random inputs before either is trusted where the reference has no
golden vector (reduce, semi-global ties).  Slow; use for small inputs only.

Written function by function from /root/reference (file:line cited on each),
keeping the reference's control flow and variable names.  Storage is a dict
per wavefront, with the reference's Lo/Hi bookkeeping.
"""

U32 = 0xFFFFFFFF
INF = 1 << 62

# wfa_backtrace_types.go:23-37
BITS, MASK = 3, 7
INS_OPEN, INS_EXT, DEL_OPEN, DEL_EXT, MISMATCH, MATCH = 1, 2, 3, 4, 5, 6
OPS = ".IIDDXMH"
ARROWS = "⊕⟼🠦↧🠧⬂⬊"


class WaveFront:
    """wfa_wavefront.go:45-183"""

    def __init__(self):
        self.Lo, self.Hi = INF, -INF
        self.c = {}

    def Set(self, k, offset, typ):              # :85-104
        self.c[k] = (offset << BITS | typ) & U32
        if k < self.Lo:
            self.Lo = k
        if k > self.Hi:
            self.Hi = k

    def Increase(self, k, delta):               # :130-150
        self.c[k] = (self.c.get(k, 0) + (delta << BITS)) & U32
        if k < self.Lo:
            self.Lo = k
        if k > self.Hi:
            self.Hi = k

    def Get(self, k):                           # :153-159
        if k < self.Lo or k > self.Hi:
            return 0, 0, False
        raw = self.c.get(k, 0)
        return raw >> BITS, raw & MASK, raw > 0

    def GetRaw(self, k):                        # :162-168
        if k < self.Lo or k > self.Hi:
            return 0, False
        raw = self.c.get(k, 0)
        return raw, raw > 0

    def Delete(self, k):                        # :171-183
        if k < self.Lo or k > self.Hi:
            return
        self.c[k] = 0
        if k == self.Hi:
            self.Hi -= 1
        elif k == self.Lo:
            self.Lo += 1


class Component:
    """wfa_component.go:37-187.  len(WaveFronts) only matters as 'bigger than
    any real score, smaller than a wrapped uint32', so a dict is enough."""

    def __init__(self):
        self.W = {}

    def HasScore(self, s):                      # :81-86
        return s in self.W

    def KRange(self, s, diff):                  # :91-101
        if diff > s:
            return 0, 0
        wf = self.W.get(s - diff)
        if wf is None:
            return 0, 0
        return wf.Lo, wf.Hi

    def Set(self, s, k, offset, typ):           # :104-115
        wf = self.W.get(s)
        if wf is None:
            wf = self.W[s] = WaveFront()
        wf.Set(k, offset, typ)

    def Get(self, s, k):                        # :142-147 (s may be a wrapped uint32)
        wf = self.W.get(s)
        if wf is None:
            return 0, 0, False
        return wf.Get(k)

    def GetRaw(self, s, k):                     # :150-155
        wf = self.W.get(s)
        if wf is None:
            return 0, False
        return wf.GetRaw(k)

    def GetAfterDiff(self, s, diff, k):         # :158-167
        if diff > s:
            return 0, 0, False
        return self.Get(s - diff, k)

    def Delete(self, s, k):                     # :182-187
        wf = self.W.get(s)
        if wf is not None:
            wf.Delete(k)


class Result:
    """wfa_cigar.go:29-214"""

    def __init__(self):
        self.Ops = []
        self.Score = 0
        self.TBegin = self.TEnd = self.QBegin = self.QEnd = 0   # reset() leaves them stale; we say 0
        self.AlignLen = self.Matches = self.Gaps = self.GapRegions = 0

    def AddN(self, op, n):                      # :118-124
        self.Ops.append((ord(op) << 32) | (n & U32))

    def process(self):                          # :136-214
        s = self.Ops
        s.reverse()
        j = 0
        pre = s[0]
        for i in range(1, len(s)):
            op = s[i]
            if op >> 32 == pre >> 32:
                pre += op & U32
                s[j] = pre
                continue
            j += 1
            if i != j:
                s[j] = s[i]
            pre = op
        del s[j + 1:]
        begin = end = 0
        for i, op in enumerate(s):
            if op >> 32 == ord('M'):
                begin = i
                break
        for i in range(len(s) - 1, -1, -1):
            if s[i] >> 32 == ord('M'):
                end = i
                break
        alen = matches = gaps = regions = 0
        for i in range(begin, end + 1):
            n = s[i] & U32
            alen += n
            o = s[i] >> 32
            if o == ord('M'):
                matches += n
            elif o in (ord('I'), ord('D')):
                gaps += n
                regions += 1
        self.AlignLen, self.Matches, self.Gaps, self.GapRegions = alen, matches, gaps, regions

    def CIGAR(self, only_aligned=False):        # :236-255 (+ trimOps :217-233)
        ops = self.Ops
        if only_aligned:
            ops = trim_ops(ops)
        return "".join("%d%s" % (op & U32, chr(op >> 32)) for op in ops)

    def AlignmentText(self, q, t, only_aligned=False):   # :259-333
        ops = self.Ops
        if only_aligned:
            q = q[self.QBegin - 1:self.QEnd]
            t = t[self.TBegin - 1:self.TEnd]
            ops = trim_ops(ops)
        Q, A, T = bytearray(), bytearray(), bytearray()
        v = h = 0
        for op in ops:
            n, o = op & U32, chr(op >> 32)
            for _ in range(n):
                if o == 'M' or o == 'X':
                    Q.append(q[v]); A.append(ord('|' if o == 'M' else ' ')); T.append(t[h]); v += 1; h += 1
                elif o == 'I':
                    Q.append(ord('-')); A.append(ord(' ')); T.append(t[h]); h += 1
                elif o in 'DH':
                    Q.append(q[v]); A.append(ord(' ')); T.append(ord('-')); v += 1
        return bytes(Q), bytes(A), bytes(T)


def trim_ops(ops):                              # wfa_cigar.go:217-233
    start = end = -1
    for i, op in enumerate(ops):
        if op >> 32 == ord('M'):
            start = i
            break
    for i in range(len(ops) - 1, -1, -1):
        if ops[i] >> 32 == ord('M'):
            end = i
            break
    return ops[start:end + 1]


class Aligner:
    """wfa.go:79-268.  ad = None or (MinWFLen, MaxDistDiff)."""

    def __init__(self, mismatch=4, gap_open=6, gap_ext=2, global_alignment=True, adaptive=None):
        self.x, self.o, self.e = mismatch, gap_open, gap_ext
        self.glob = global_alignment
        self.ad = adaptive
        self.M = self.I = self.D = None

    # wfa.go:143-184
    def initComponents(self, q, t):
        self.M, self.I, self.D = Component(), Component(), Component()
        m, n = len(t), len(q)
        M = self.M
        if q[0] == t[0]:
            M.Set(0, 0, 1, MATCH)
        else:
            M.Set(self.x, 0, 1, MISMATCH)
        if not self.glob:
            for k in range(1, m):
                if q[0] == t[k]:
                    M.Set(0, k, k + 1, MATCH)
                else:
                    M.Set(self.x, k, k + 1, MISMATCH)
            for k in range(1, n):
                if q[k] == t[0]:
                    M.Set(0, -k, 1, MATCH)
                else:
                    M.Set(self.x, -k, 1, MISMATCH)

    # wfa.go:201-268
    def Align(self, q, t):
        q, t = bytes(q), bytes(t)
        m, n = len(t), len(q)
        if n == 0 or m == 0:
            raise ValueError("wfa: invalid empty sequence")
        self.initComponents(q, t)
        Ak = m - n
        Aoffset = m
        M = self.M
        s = 0
        reduce = self.ad is not None
        minWFLen = self.ad[0] if reduce else 0
        while True:
            if M.HasScore(s):
                lo, hi = self.extend(q, t, s)
                offset, _, _ = M.GetAfterDiff(s, 0, Ak)
                if offset >= Aoffset:
                    break
                if reduce and hi - lo + 1 >= minWFLen:
                    self.reduce(q, t, s)
            s += 1
            self.next(q, t, s)
        self.final_score = s
        minS, lastK = s, Ak
        if not self.glob:
            minS, lastK = self.backtraceStartPosistion(q, t, s)
        return self.backTrace(q, t, minS, lastK)

    # wfa.go:270-375
    def backtraceStartPosistion(self, q, t, s):
        M = self.M
        m, n = len(t), len(q)
        minS = s
        Ak = m - n
        lastK = Ak
        _s = s
        while True:
            if not M.HasScore(_s):
                if _s == 0:
                    break
                _s -= 1
                continue
            lo, hi = M.KRange(_s, 0)
            lastRowOrCol = False
            k = Ak
            while True:
                if k < lo:
                    break
                offset, _, ok = M.GetAfterDiff(_s, 0, k)
                if not ok:
                    k -= 1
                    continue
                h = offset
                v = h - k
                if v <= 0 or v > n or h > m:
                    break
                if (v == n and h >= n) or (h == m and v >= m):
                    lastRowOrCol = True
                    break
                k -= 1
            if lastRowOrCol and _s <= minS:
                lastK = k
                minS = _s
            lastRowOrCol = False
            k = Ak + 1
            while True:
                if k > hi:
                    break
                offset, _, ok = M.GetAfterDiff(_s, 0, k)
                if not ok:
                    k += 1
                    continue
                h = offset
                v = h - k
                if v <= 0 or v > n or h > m:
                    break
                if (v == n and h >= n) or (h == m and v >= m):
                    lastRowOrCol = True
                    break
                k += 1
            if lastRowOrCol and _s <= minS:
                lastK = k
                minS = _s
            if _s == 0:
                break
            _s -= 1
        return minS, lastK

    # wfa.go:381-458 -- kept in the reference's two-stage shape (8-byte blocks, then bytes)
    def extend(self, q, t, s):
        wf = self.M.W[s]
        lo, hi = wf.Lo, wf.Hi
        lenQ, lenT = len(q), len(t)
        for k in range(hi, lo - 1, -1):
            offset, _, ok = wf.Get(k)
            if not ok:
                continue
            h = offset
            v = h - k
            if v <= 0 or v >= lenQ or h >= lenT:
                continue
            if v + 8 <= lenQ and h + 8 <= lenT:
                N = 0
                while True:
                    q8 = int.from_bytes(q[v:v + 8], "big")
                    t8 = int.from_bytes(t[h:h + 8], "big")
                    x = q8 ^ t8
                    n = (64 - x.bit_length()) >> 3
                    v += n
                    h += n
                    N += n
                    if n < 8 or v + 8 >= lenQ or h + 8 >= lenT:
                        break
                if N == 0:
                    continue
                wf.Increase(k, N)
                if not (n == 8 and v < lenQ and h < lenT):
                    continue
            N = 0
            while q[v] == t[h]:
                v += 1
                h += 1
                N += 1
                if v == lenQ or h == lenT:
                    break
            if N == 0:
                continue
            wf.Increase(k, N)
        return lo, hi

    # wfa.go:461-540
    def reduce(self, q, t, s):
        wf = self.M.W[s]
        lo, hi = wf.Lo, wf.Hi
        lenQ, lenT = len(q), len(t)
        ds = []
        minDist = INF
        for k in range(lo, hi + 1):
            offset, _, ok = wf.Get(k)
            if not ok:
                ds.append(-1)
                continue
            h = offset
            v = h - k
            if v < 0 or v >= lenQ or h >= lenT:
                ds.append(-1)
                continue
            d = max(lenT - h, lenQ - v)
            ds.append(d)
            if d < minDist:
                minDist = d
        _lo, _hi = lo, hi
        maxDistDiff = self.ad[1]
        updateLo = True
        found = False
        for i, d in enumerate(ds):
            if d < 0:
                continue
            if d - minDist > maxDistDiff:
                found = True
                if updateLo:
                    _lo = lo + i + 1
                ds[i] = -1
            else:
                updateLo = False
        if found:
            for i in range(len(ds) - 1, -1, -1):
                if ds[i] >= 0:
                    _hi = lo + i
                    break
        for k in range(lo, _lo):
            wf.Delete(k)
            self.I.Delete(s, k)
            self.D.Delete(s, k)
        for k in range(_hi + 1, hi + 1):
            wf.Delete(k)
            self.I.Delete(s, k)
            self.D.Delete(s, k)
        wf.Lo, wf.Hi = _lo, _hi

    # wfa.go:549-700
    def next(self, q, t, s):
        M, I, D = self.M, self.I, self.D
        x, oe, e = self.x, self.o + self.e, self.e
        lenQ, lenT = len(q), len(t)
        loMismatch, hiMismatch = M.KRange(s, x)
        loGapOpen, hiGapOpen = M.KRange(s, oe)
        loInsert, hiInsert = I.KRange(s, e)
        loDelete, hiDelete = D.KRange(s, e)
        hi = min(lenT - 1, max(hiMismatch, hiGapOpen, hiInsert, hiDelete) + 1)
        lo = max(-(lenQ - 1), min(loMismatch, loGapOpen, loInsert, loDelete) - 1)
        for k in range(lo, hi + 1):
            updatedI = updatedD = False
            wfaTypeI = wfaTypeD = wfaTypeM = 0
            v1, _, fromM = M.GetAfterDiff(s, oe, k - 1)
            v2, _, fromI = I.GetAfterDiff(s, e, k - 1)
            if fromM and v1 > lenT:
                fromM, v1 = False, 0
            if fromI and v2 > lenT:
                fromI, v2 = False, 0
            Isk = max(v1, v2) + 1
            if fromM or fromI:
                if fromM and fromI:
                    wfaTypeI = INS_OPEN if v1 >= v2 else INS_EXT
                elif fromM:
                    wfaTypeI = INS_OPEN
                else:
                    wfaTypeI = INS_EXT
                updatedI = True
                I.Set(s, k, Isk, wfaTypeI)
            else:
                Isk = 0

            v1, _, fromM = M.GetAfterDiff(s, oe, k + 1)
            v2, _, fromD = D.GetAfterDiff(s, e, k + 1)
            if fromM and v1 - k > lenQ:
                fromM, v1 = False, 0
            if fromD and v2 - k > lenQ:
                fromD, v2 = False, 0
            Dsk = max(v1, v2)
            if fromM or fromD:
                if fromM and fromD:
                    wfaTypeD = DEL_OPEN if v1 >= v2 else DEL_EXT
                elif fromM:
                    wfaTypeD = DEL_OPEN
                else:
                    wfaTypeD = DEL_EXT
                updatedD = True
                D.Set(s, k, Dsk, wfaTypeD)
            else:
                Dsk = 0

            v1, _, fromM = M.GetAfterDiff(s, x, k)
            if fromM and (v1 > lenT or v1 - k > lenQ):
                fromM, v1 = False, 0
            Msk = max(Isk, Dsk, v1 + 1)
            if updatedI or updatedD or fromM:
                if updatedI and updatedD and fromM:
                    if Msk == v1 + 1:
                        wfaTypeM = MISMATCH
                    elif Msk == Isk:
                        wfaTypeM = wfaTypeI
                    else:
                        wfaTypeM = wfaTypeD
                elif updatedI:
                    if updatedD:
                        wfaTypeM = wfaTypeI if Msk == Isk else wfaTypeD
                    elif fromM:
                        wfaTypeM = MISMATCH if Msk == v1 + 1 else wfaTypeI
                    else:
                        wfaTypeM = wfaTypeI
                elif updatedD:
                    if fromM:
                        wfaTypeM = MISMATCH if Msk == v1 + 1 else wfaTypeD
                    else:
                        wfaTypeM = wfaTypeD
                else:
                    wfaTypeM = MISMATCH
                M.Set(s, k, Msk, wfaTypeM)

    # wfa.go:703-983
    def backTrace(self, q, t, s, Ak):
        semiGlobal = not self.glob
        M, I, D = self.M, self.I, self.D
        x, o, e = self.x, self.o, self.e
        lenQ, lenT = len(q), len(t)
        cigar = Result()
        cigar.Score = s
        qBegin = tBegin = 0
        offset0 = Isk = Dsk = 0
        fromItself = False
        M0 = None
        k = Ak
        firstMatch = True
        offset, _ = M.GetRaw(s, k)
        previousFromM = True
        wfaType = offset & MASK
        h = offset >> BITS
        v = h - k
        if h < lenT:
            cigar.AddN(OPS[INS_OPEN], lenT - h)
        elif v < lenQ:
            cigar.AddN('H', lenQ - v)

        while v > 0 and h > 0:
            sMismatch = (s - x) & U32
            sGapOpen = (s - o - e) & U32
            sGapExt = (s - e) & U32
            fromMI = fromMD = False
            if wfaType == INS_EXT:
                v1, _, fromM = M.Get(sGapOpen, k - 1)
                v2, _, fromI = I.Get(sGapExt, k - 1)
                if fromM or fromI:
                    fromMI = True
                    offset0 = max(v1, v2) + 1
                else:
                    offset0 = 0
                M0 = I
            elif wfaType == DEL_EXT:
                v1, _, fromM = M.Get(sGapOpen, k + 1)
                v2, _, fromD = D.Get(sGapExt, k + 1)
                if fromM or fromD:
                    fromMD = True
                    offset0 = max(v1, v2)
                else:
                    offset0 = 0
                M0 = D
            else:
                v1, _, fromM = M.Get(sGapOpen, k - 1)
                v2, _, fromI = I.Get(sGapExt, k - 1)
                if fromM or fromI:
                    fromMI = True
                    Isk = max(v1, v2) + 1
                else:
                    Isk = 0
                v1, _, fromM = M.Get(sGapOpen, k + 1)
                v2, _, fromD = D.Get(sGapExt, k + 1)
                if fromM or fromD:
                    fromMD = True
                    Dsk = max(v1, v2)
                else:
                    Dsk = 0
                v1, _, fromM = M.Get(sMismatch, k)
                if fromMI or fromMD or fromM:
                    offset0 = max(Isk, Dsk, v1 + 1)
                    fromItself = False
                else:
                    fromItself = True
                M0 = M
            if fromItself:
                break
            if offset0 == 0:
                break
            h0 = offset0
            if previousFromM:
                nMatches = h - h0
                if nMatches > 0:
                    if firstMatch:
                        firstMatch = False
                        cigar.TEnd, cigar.QEnd = h, v
                    cigar.AddN(OPS[MATCH], nMatches)
                offset = offset0
                h = offset
                v = h - k
                if wfaType == MATCH:
                    tBegin, qBegin = h, v
                elif nMatches > 0:
                    tBegin, qBegin = h + 1, v + 1
                if h <= 0 or v <= 0:
                    break
            cigar.AddN(OPS[wfaType], 1)
            if semiGlobal and (h == 1 or v == 1):
                break
            previousFromM = True
            if wfaType == MISMATCH:
                s = sMismatch
                h -= 1
            elif wfaType == INS_OPEN:
                s = sGapOpen
                k -= 1
                h -= 1
            elif wfaType == INS_EXT:
                s = sGapExt
                k -= 1
                h -= 1
                previousFromM = False
            elif wfaType == DEL_OPEN:
                s = sGapOpen
                k += 1
            elif wfaType == DEL_EXT:
                s = sGapExt
                k += 1
                previousFromM = False
            else:
                break
            v = h - k
            offset, ok = M0.GetRaw(s, k)
            if not ok:
                break
            wfaType = offset & MASK

        if h > 0 and v > 0:
            nMatches = min(h, v) - 1
            if nMatches > 0:
                if firstMatch:
                    firstMatch = False
                    cigar.TEnd, cigar.QEnd = h, v
                cigar.AddN(OPS[MATCH], nMatches)
                h -= nMatches
                v -= nMatches
                if wfaType == MATCH:
                    tBegin, qBegin = h, v
                elif nMatches > 0:
                    tBegin, qBegin = h + 1, v + 1
            elif wfaType == MATCH:
                tBegin, qBegin = h, v
                if firstMatch:
                    firstMatch = False
                    cigar.TEnd, cigar.QEnd = h, v
            cigar.AddN(OPS[wfaType], 1)
        if v > 1:
            cigar.AddN('H', v - 1)
        if h > 1:
            cigar.AddN(OPS[INS_OPEN], h - 1)
        cigar.TBegin, cigar.QBegin = tBegin, qBegin
        cigar.process()
        return cigar

    # wfa_component_plot.go:41-209, returned as a matrix of (score, type) or None
    def plot_matrix(self, q, t, comp="M", notChangeToMatch=True, maxScore=-1):
        M, I, D = self.M, self.I, self.D
        _M = {"M": M, "I": I, "D": D}[comp]
        x, oe, e = self.x, self.o + self.e, self.e
        lenQ, lenT = len(q), len(t)
        isM = True      # the reference tests algn.M.IsM, which is always true (wfa.go:97)
        mat = [[None] * lenT for _ in range(lenQ)]
        vp = hp = 0     # function-scoped in the reference (:60), so stale values carry over
        for s in sorted(_M.W):
            wf = _M.W[s]
            if maxScore >= 0 and s > maxScore:
                break
            for k in range(wf.Lo, wf.Hi + 1):
                offset, wfaType, ok = wf.Get(k)
                if not ok:
                    continue
                h = offset - 1
                v = h - k
                if v < 0 or h < 0 or v >= lenQ or h >= lenT:
                    continue
                if mat[v][h] is not None:
                    continue
                mat[v][h] = (s, wfaType)
                if not isM or q[v] != t[h]:
                    continue
                if wfaType == INS_EXT:
                    v1 = M.GetAfterDiff(s, oe, k - 1)[0]
                    v2 = I.GetAfterDiff(s, e, k - 1)[0]
                    offset0 = max(v1, v2) + 1
                elif wfaType == DEL_EXT:
                    v1 = M.GetAfterDiff(s, oe, k + 1)[0]
                    v2 = D.GetAfterDiff(s, e, k + 1)[0]
                    offset0 = max(v1, v2)
                else:
                    v1 = M.GetAfterDiff(s, oe, k - 1)[0]
                    v2 = I.GetAfterDiff(s, e, k - 1)[0]
                    Isk = max(v1, v2) + 1
                    v1 = M.GetAfterDiff(s, oe, k + 1)[0]
                    v2 = D.GetAfterDiff(s, e, k + 1)[0]
                    Dsk = max(v1, v2)
                    v1 = M.GetAfterDiff(s, x, k)[0]
                    offset0 = max(Isk, Dsk, v1 + 1)
                h00 = offset0 - 1
                if h == h00:
                    continue
                v0, h0 = v, h
                if not notChangeToMatch:
                    mat[v0][h0] = (s, MATCH)
                n = 0
                while True:
                    h -= 1
                    v -= 1
                    if v < 0 or h < 0:
                        break
                    n += 1
                    if mat[v][h] is not None:
                        continue
                    mat[v][h] = (s, MATCH) if not notChangeToMatch else (s, wfaType)
                    vp, hp = v, h
                    if q[v] != t[h] or h == h00:
                        break
                if n == 0:
                    vp, hp = v0, h0
                if not notChangeToMatch:
                    mat[vp][hp] = (s, wfaType)
        return mat
