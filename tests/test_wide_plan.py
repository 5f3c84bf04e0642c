"""Host logic of the WIDE worker that needs no GPU: the launch geometry (wfacuda_wide_plan) -- cluster size,
diagonals per CTA, block size, shared memory -- for the widths the worker is meant for and beyond."""
import pytest

from wfa_b200 import api

SMEM = 232448            # opt-in shared memory per CTA on sm_100a


def entries(n, m):
    return (n + 15) // 16 + (m + 15) // 16 + 2


def test_config4_runs_on_two_cta_clusters(built_lib):
    # 10 kbp reads (5 % indels) in 12 kbp windows: n + m - 1 diagonals, all of them live from score 0 on (semi-global)
    for n in (9900, 10000, 10100):
        c, seg, th, sm = api.wide_plan(n + 12000 - 1, entries(n, 12000), SMEM)
        assert c == 2 and th == 1024 and seg % 64 == 0 and c * seg >= n + 12000 - 1 and sm <= SMEM - 1024
        assert sm == 1280 + 9 * (seg + 4) * 2 + 8 + entries(n, 12000) * 8         # head, 9 rows of 16-bit offsets, windows


@pytest.mark.parametrize("w,ent", [(1, 4), (63, 10), (64, 10), (65, 12), (1000, 66), (11000, 700), (11300, 700), (11400, 700),
                                    (22000, 1400), (40000, 2600), (60000, 3800), (80000, 3000)])
def test_plan_properties(built_lib, w, ent):
    c, seg, th, sm = api.wide_plan(w, ent, SMEM)
    assert c in (1, 2, 4, 8) and c * seg >= w and seg % 64 == 0 and sm + 1344 <= SMEM
    assert 128 <= th <= 1024 and th % 32 == 0 and (th == 1024 or th * 2 >= seg)
    if c > 1:                                                   # the smallest cluster that fits: half as many CTAs would not
        half = ((w + c // 2 - 1) // (c // 2) + 63) // 64 * 64
        assert 1280 + 9 * (half + 4) * 2 + 8 + ent * 8 + 1344 > SMEM


def test_too_wide_for_eight_ctas(built_lib):
    assert api.wide_plan(120000, 7600, SMEM) is None and api.wide_plan(85000, 5400, SMEM) is None      # such pairs go to the CTA worker
    assert api.wide_plan(1000, 66, 16384) is not None and api.wide_plan(60000, 3800, 49152) is None
