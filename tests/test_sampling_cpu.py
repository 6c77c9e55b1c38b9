"""On-device filtered negative sampler (jmac_b200/sampling.py; modules/load/data_loader.py:36-47 for a whole batch).
Pure torch on integer tensors, so its logic is checked on the CPU; distributional parity with the reference's
np.random.choice(all_ent[mask], k, replace=False)."""
import numpy as np
import pytest
import torch

from jmac_b200.sampling import TripleIndex, sample_negatives_torch as sample_negatives


def make_index(n_ent=50, n_rel=4, n_tri=300, seed=0):
    rng = np.random.default_rng(seed)
    tri = np.stack([rng.integers(0, n_ent, n_tri), rng.integers(0, n_rel, n_tri), rng.integers(0, n_ent, n_tri)], 1)
    return tri, TripleIndex(tri, n_ent, n_rel)


def test_negatives_are_valid_distinct_and_seeded():
    tri, idx = make_index()
    true = {}
    for h, r, t in tri:
        true.setdefault((int(h), int(r)), set()).add(int(t))
    h, r = torch.from_numpy(tri[:, 0]), torch.from_numpy(tri[:, 1])
    g = torch.Generator().manual_seed(3)
    neg = sample_negatives(idx, h, r, 25, generator=g)
    assert neg.shape == (len(tri), 25) and neg.dtype == torch.int64
    for b in range(len(tri)):
        row = neg[b].tolist()
        assert len(set(row)) == 25 and all(0 <= v < 50 for v in row)
        assert not (set(row) & true[(int(h[b]), int(r[b]))])
    again = sample_negatives(idx, h, r, 25, generator=torch.Generator().manual_seed(3))
    assert torch.equal(neg, again)
    assert torch.equal(idx.n_true_tails(h, r), torch.tensor([len(true[(int(a), int(b))]) for a, b in zip(h, r)]))


def test_nearly_exhausted_rows_and_errors():
    n_ent = 12
    tri = np.array([[0, 0, t] for t in range(9)] + [[1, 0, 5]])            # (0,0) has 9 true tails: only 3 valid entities
    idx = TripleIndex(tri, n_ent, 2)
    neg = sample_negatives(idx, torch.tensor([0, 1, 0]), torch.tensor([0, 0, 0]), 3, generator=torch.Generator().manual_seed(0))
    assert sorted(neg[0].tolist()) == [9, 10, 11] and sorted(neg[2].tolist()) == [9, 10, 11]
    assert 5 not in neg[1].tolist() and len(set(neg[1].tolist())) == 3
    with pytest.raises(ValueError):                                        # np.random.choice(..., replace=False) raises there too
        sample_negatives(idx, torch.tensor([0]), torch.tensor([0]), 4)
    assert sample_negatives(idx, torch.tensor([], dtype=torch.long), torch.tensor([], dtype=torch.long), 3).shape == (0, 3)
    with pytest.raises(ValueError):
        TripleIndex(np.array([[0, 0, 99]]), n_ent, 2)


def test_distribution_is_uniform_over_the_valid_entities():
    """First position and membership frequencies against the reference's sampler distribution (uniform without
    replacement over the non-tails): chi-square well inside the 99.9 % quantile."""
    n_ent, k, B = 20, 5, 40000
    tri = np.array([[0, 0, t] for t in (1, 3, 5, 7)])
    idx = TripleIndex(tri, n_ent, 1)
    neg = sample_negatives(idx, torch.zeros(B, dtype=torch.long), torch.zeros(B, dtype=torch.long), k, generator=torch.Generator().manual_seed(11))
    valid = [e for e in range(n_ent) if e not in (1, 3, 5, 7)]
    first = np.bincount(neg[:, 0].numpy(), minlength=n_ent)[valid]
    member = np.bincount(neg.reshape(-1).numpy(), minlength=n_ent)[valid]
    assert np.bincount(neg.reshape(-1).numpy(), minlength=n_ent)[[1, 3, 5, 7]].sum() == 0
    for counts, total in ((first, B), (member, B * k)):
        exp = total / len(valid)
        chi2 = float(((counts - exp) ** 2 / exp).sum())
        assert chi2 < 45.0, chi2          # 15 degrees of freedom: P(chi2 > 37.7) = 1e-3 (membership counts are negatively correlated: conservative)
    # the reference's sampler on the same problem, same statistic
    rng = np.random.default_rng(0)
    ref = np.stack([rng.choice(np.array(valid), k, replace=False) for _ in range(4000)])
    ref_first = np.bincount(ref[:, 0], minlength=n_ent)[valid]
    assert abs(ref_first.mean() - 4000 / len(valid)) < 1e-9
