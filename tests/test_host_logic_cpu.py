"""Host-side helpers that need no GPU: destination blocks of comp_op='mult', the padded layer width."""
import numpy as np
import torch


def test_mult_block_bounds_respect_the_edge_budget():
    from jmac_b200.model import mult_block_bounds
    rng = np.random.default_rng(0)
    for n, e_max in [(1, 5), (50, 7), (1000, 64), (1000, 10_000_000)]:
        deg = torch.from_numpy(rng.integers(0, 12, n))
        deg[n // 3] = 500                                        # a hub above most budgets
        b = mult_block_bounds(deg, e_max)
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:]))
        for lo, hi in zip(b, b[1:]):
            assert int(deg[lo:hi].sum()) <= e_max or hi - lo == 1    # over budget only as a single (hub) node
        if e_max >= int(deg.sum()):
            assert b == [0, n]
    assert mult_block_bounds(torch.zeros(4, dtype=torch.int64), 3) == [0, 4]      # no edges at all: one block


def test_padded_width():
    from jmac_b200.functional import padded_width
    assert [padded_width(d) for d in (32, 256, 300, 200, 33, 512)] == [32, 256, 320, 224, 64, 512]
