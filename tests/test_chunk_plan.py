"""Host logic of the pipelined wfacuda_align_batch that needs no GPU: the chunk boundaries
(wfacuda_chunk_plan) -- complete, ordered, whole LANE groups, small first and last chunks."""
import numpy as np
import pytest

from wfa_b200 import api


@pytest.mark.parametrize("n,c,levels", [(1_000_000, 66_674, 2), (340_000, 53_333, 2), (1_000_003, 66_674, 3),
                                         (2_000_000, 100_000, 1), (399_999, 53_333, 0), (1_000_000, 66_674, -1),
                                         (150_000, 53_333, 2), (70_000, 32_768, 2), (5, 66_674, 2), (0, 66_674, 2),
                                         (1_250, 320, 2), (10_000, 2_528, 2), (256, 64, 2)])      # a few hundred long pairs: four chunks side by side
def test_chunk_plan_properties(built_lib, n, c, levels):
    cuts = api.chunk_plan(n, c, levels).astype(np.int64)
    assert cuts[0] == 0 and cuts[-1] == n
    sizes = np.diff(cuts)
    if n == 0:
        assert len(sizes) == 0
        return
    assert (sizes > 0).all()                                   # ordered, no empty chunk
    assert (cuts[1:-1] % 32 == 0).all()                        # inner boundaries keep whole groups of 32 pairs
    assert sizes.max() <= c + 64                               # no chunk much above the target size
    graded = levels >= 0 and n >= 6 * c and c >= 32768
    if graded:
        assert list(sizes[:3]) == [(c // 8) & ~31, (c // 4) & ~31, (c // 2) & ~31]          # C/8, C/4, C/2 first
        for l in range(1, levels + 1):                                                     # ... C/4, C/2 at the end
            want = (c >> (levels - l + 1)) & ~31
            got = sizes[len(sizes) - l]
            assert abs(int(got) - want) < 32 or l == 1, (l, got, want)                     # the very last chunk takes the remainder
        mid = sizes[3:len(sizes) - levels]
        assert mid.max() - mid.min() <= 64                                                  # equal middle chunks
    else:
        assert sizes.max() - sizes.min() <= 64 or len(sizes) == 1
    if c < 32768 and n >= 256:
        assert len(sizes) == 4, sizes                              # the long-pair rule of wfacuda_align_batch: chunk = ceil(n / 4) rounded up to 32
