"""A third, independent check of the oracle's scores: for global alignment without heuristic the
wavefront score must be the optimal gap-affine cost of the classic three-matrix dynamic programme
(Gotoh) under the reference's boundary condition -- initComponents (wfa.go:155-158) starts every
alignment by pairing q[0] with t[0], so the first column of the path is a match or a mismatch,
never a gap.  Random small pairs over several penalty sets; the DP shares no code with either
restatement."""
import random

import pytest

import oracle_lib

INF = 10 ** 9


def dp_cost(q, t, x, o, e):
    n, m = len(q), len(t)
    H = [[INF] * (m + 1) for _ in range(n + 1)]
    I = [[INF] * (m + 1) for _ in range(n + 1)]     # gap in the query: consumes a target base
    D = [[INF] * (m + 1) for _ in range(n + 1)]     # gap in the target: consumes a query base
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            if i == 1 and j == 1:
                H[1][1] = 0 if q[0] == t[0] else x          # forced diagonal start
                continue
            I[i][j] = min(H[i][j - 1] + o + e, I[i][j - 1] + e)
            D[i][j] = min(H[i - 1][j] + o + e, D[i - 1][j] + e)
            H[i][j] = min(H[i - 1][j - 1] + (0 if q[i - 1] == t[j - 1] else x), I[i][j], D[i][j])
    return H[n][m]


@pytest.mark.parametrize("pen", [(4, 6, 2), (1, 0, 1), (3, 1, 2), (2, 3, 1), (5, 2, 3), (2, 12, 2)])
def test_score_is_the_optimal_gap_affine_cost(pen):
    x, o, e = pen
    rng = random.Random(sum(pen))
    orc = oracle_lib.Oracle(mismatch=x, gap_open=o, gap_ext=e)
    for it in range(250):
        alpha = b"ACGT" if it % 3 else b"AC"
        n = rng.randint(1, 40)
        q = bytes(rng.choice(alpha) for _ in range(n))
        if it % 2:
            t = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 40)))
        else:
            t = bytearray(q)
            for _ in range(rng.randint(0, 6)):
                j = rng.randrange(len(t) + 1)
                r = rng.random()
                if r < 0.4 and j < len(t):
                    t[j] = rng.choice(alpha)
                elif r < 0.7:
                    t.insert(j, rng.choice(alpha))
                elif j < len(t) and len(t) > 1:
                    del t[j]
            t = bytes(t)
        r = orc.align(q, t)
        assert r["status"] == 0
        assert r["score"] == dp_cost(q, t, x, o, e), (pen, q, t, oracle_lib.ops_to_cigar(r["ops"]))
    orc.close()


def dp_cost_semiglobal(q, t, x, o, e):
    """Semi-global as the reference defines it: the path may start at any cell of the first row or
    column (initComponents seeds them all, wfa.go:160-183 -- still a match or mismatch, never a gap)
    and ends on the last row or column, but only where backtraceStartPosistion accepts a hit
    (wfa.go:306-323, :341-358): v == n with h >= n, or h == m with v >= m."""
    n, m = len(q), len(t)
    H = [[INF] * (m + 1) for _ in range(n + 1)]
    I = [[INF] * (m + 1) for _ in range(n + 1)]
    D = [[INF] * (m + 1) for _ in range(n + 1)]
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            sub = 0 if q[i - 1] == t[j - 1] else x
            start = sub if (i == 1 or j == 1) else INF
            diag = H[i - 1][j - 1] + sub if (i > 1 and j > 1) else INF
            I[i][j] = min(H[i][j - 1] + o + e, I[i][j - 1] + e)
            D[i][j] = min(H[i - 1][j] + o + e, D[i - 1][j] + e)
            H[i][j] = min(start, diag, I[i][j], D[i][j])
    ends = [H[n][j] for j in range(1, m + 1) if j >= n] + [H[i][m] for i in range(1, n + 1) if i >= m]
    return min(ends)


@pytest.mark.parametrize("pen", [(4, 6, 2), (1, 0, 1), (3, 1, 2), (2, 3, 1), (5, 2, 3)])
def test_semiglobal_score_is_the_optimal_cost_over_the_reference_s_start_and_end_cells(pen):
    x, o, e = pen
    rng = random.Random(100 + sum(pen))
    orc = oracle_lib.Oracle(mismatch=x, gap_open=o, gap_ext=e, global_alignment=False)
    for it in range(250):
        n = rng.randint(1, 30)
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        if it % 2:
            t = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 36)))
        else:
            t = bytearray(q)
            for _ in range(rng.randint(0, 4)):
                j = rng.randrange(len(t) + 1)
                r = rng.random()
                if r < 0.4 and j < len(t):
                    t[j] = rng.choice(b"ACGT")
                elif r < 0.7:
                    t.insert(j, rng.choice(b"ACGT"))
                elif j < len(t) and len(t) > 1:
                    del t[j]
            t = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 8))) + bytes(t) + bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 8)))
        r = orc.align(q, t)
        assert r["status"] == 0
        assert r["score"] == dp_cost_semiglobal(q, t, x, o, e), (pen, q, t, oracle_lib.ops_to_cigar(r["ops"]))
    orc.close()
