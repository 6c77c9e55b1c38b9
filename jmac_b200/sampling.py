"""On-device filtered negative sampling for the completion scorer.

The reference draws, per training triple (h, r, t), ``neg_num`` DISTINCT entities uniformly from the entities that are
not a known tail of (h, r): ``np.random.choice(all_ent[mask], neg_num, replace=False)`` in
``TrainDataset.get_neg_sample`` (modules/load/data_loader.py:36-47) -- an O(num_ent) mask per triple on the CPU, which
is why train.py runs it in 12 DataLoader workers (train.py:288-294).  Here a whole batch is sampled at once on the
device the indices live on: candidates are drawn i.i.d. uniform, candidates that are true tails or repeat an earlier
candidate of the same row are rejected, and the first ``k`` survivors of every row are kept.  Sequential rejection of
an i.i.d. uniform stream yields exactly the uniform distribution over ordered k-subsets of the valid entities, i.e.
the reference's distribution; the random STREAM differs (Philox instead of numpy's global generator), so
parity is distributional -- the parity tests of the scorer feed precomputed negatives (north_star).

``sample_negatives`` is the CUDA kernel (csrc/sampling.cu: one thread per row, Philox-4x32-10 keyed by (seed, row), binary
search in the sorted triple keys; one launch per batch, no host round trip besides the k-feasibility check).
``sample_negatives_torch`` states the same procedure in batched torch ops; it is the host-side reference the CPU tests
(validity, exhaustion, chi-square uniformity) exercise, and what the GPU test compares the kernel's distribution with.
"""
from typing import Optional

import numpy as np
import torch


class TripleIndex:
    """Sorted int64 keys ``(h * n_rel + r) * n_ent + t`` of the known triples: membership of (h, r, candidate) is one
    ``searchsorted``.  The reference keeps ``triple_dict[(h, r)] -> list of tails`` (train.py:262-279)."""

    def __init__(self, triples, n_ent: int, n_rel: int, device=None):
        t = torch.as_tensor(np.asarray(triples)) if not torch.is_tensor(triples) else triples
        t = t.to(device=device, dtype=torch.int64).reshape(-1, 3)
        if int(n_ent) * int(n_rel) * int(n_ent) >= 2 ** 62:
            raise ValueError("key (h * n_rel + r) * n_ent + t does not fit 62 bits")
        self.n_ent, self.n_rel = int(n_ent), int(n_rel)
        if t.numel() and (int(t[:, [0, 2]].min()) < 0 or int(t[:, [0, 2]].max()) >= n_ent or int(t[:, 1].min()) < 0 or int(t[:, 1].max()) >= n_rel):
            raise ValueError("triple entry out of range")
        self.keys = torch.unique((t[:, 0] * self.n_rel + t[:, 1]) * self.n_ent + t[:, 2])      # sorted
        hr = torch.div(self.keys, self.n_ent, rounding_mode="floor")
        self._hr, self._cnt = torch.unique_consecutive(hr, return_counts=True)

    @property
    def device(self):
        return self.keys.device

    def contains(self, h: torch.Tensor, r: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Elementwise: is (h, r, t) a known triple?  (broadcasting shapes)"""
        key = (h * self.n_rel + r) * self.n_ent + t
        if self.keys.numel() == 0:
            return torch.zeros_like(key, dtype=torch.bool)
        pos = torch.searchsorted(self.keys, key.reshape(-1)).clamp_(max=self.keys.numel() - 1)
        return (self.keys[pos] == key.reshape(-1)).reshape(key.shape)

    def n_true_tails(self, h: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
        """Number of known tails of every (h, r)."""
        hr = h * self.n_rel + r
        if self._hr.numel() == 0:
            return torch.zeros_like(hr)
        pos = torch.searchsorted(self._hr, hr).clamp_(max=self._hr.numel() - 1)
        return torch.where(self._hr[pos] == hr, self._cnt[pos], torch.zeros_like(hr))


def _first_occurrence(cand: torch.Tensor) -> torch.Tensor:
    """[B, C] bool: True where a value does not appear earlier in its row."""
    B, C = cand.shape
    order = torch.argsort(cand, dim=1, stable=True)                     # equal values keep their stream order
    sv = torch.gather(cand, 1, order)
    first_sorted = torch.ones_like(sv, dtype=torch.bool)
    first_sorted[:, 1:] = sv[:, 1:] != sv[:, :-1]
    out = torch.empty_like(first_sorted)
    out.scatter_(1, order, first_sorted)
    return out


def sample_negatives(index: TripleIndex, h, r, k: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """[B, k] int64 on the device: for every (h[b], r[b]), k distinct entities drawn uniformly from those that are not known
    tails of (h[b], r[b]) -- ``TrainDataset.get_neg_sample`` for a whole batch, by the CUDA kernel ``jmac_sample_negatives``.
    The stream is keyed by a 62-bit seed drawn from ``generator`` (default: torch's global CPU generator, so
    ``torch.manual_seed`` makes it reproducible).  Raises ValueError if some (h, r) has fewer than k valid entities (the
    reference's ``np.random.choice(..., replace=False)`` raises there too)."""
    from . import _lib
    from ._lib import check, ptr, stream_ptr
    dev = index.device
    if dev.type != "cuda":
        raise _lib.JmacError("sample_negatives is a CUDA kernel; sample_negatives_torch is the host-side reference")
    lib = _lib.load()
    h = torch.as_tensor(h, device=dev).to(torch.int64).reshape(-1).contiguous()
    r = torch.as_tensor(r, device=dev).to(torch.int64).reshape(-1).contiguous()
    B, N = h.numel(), index.n_ent
    out = torch.empty(B, k, dtype=torch.int64, device=dev)
    if B == 0 or k == 0:
        return out
    if bool((N - index.n_true_tails(h, r) < k).any()):
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=generator.device if generator is not None else "cpu").item())
    _lib.raise_if_flagged()
    keys = index.keys.contiguous()
    with torch.cuda.device(dev):
        check(lib.jmac_sample_negatives(ptr(keys), int(keys.numel()), ptr(h), ptr(r), B, N, index.n_rel, int(k), seed, ptr(out), stream_ptr()),
              "jmac_sample_negatives")
    return out


def sample_negatives_torch(index: TripleIndex, h, r, k: int, generator: Optional[torch.Generator] = None, oversample: float = 1.5) -> torch.Tensor:
    """The same procedure in batched torch ops, wherever ``index`` lives (host-side reference, see the module docstring)."""
    dev = index.device
    h = torch.as_tensor(h, device=dev).to(torch.int64).reshape(-1)
    r = torch.as_tensor(r, device=dev).to(torch.int64).reshape(-1)
    B, N = h.numel(), index.n_ent
    out = torch.empty(B, k, dtype=torch.int64, device=dev)
    if B == 0 or k == 0:
        return out
    if bool((N - index.n_true_tails(h, r) < k).any()):
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    rows = torch.arange(B, device=dev)
    C = max(k + 8, int(k * oversample))
    kept = torch.empty(B, 0, dtype=torch.int64, device=dev)              # survivors so far, stream order (all rows)
    for _ in range(64):
        hr, rr = h[rows], r[rows]
        cand = torch.randint(0, N, (rows.numel(), C), device=dev, generator=generator)
        stream = torch.cat([kept[rows] if kept.shape[1] else kept.new_empty(rows.numel(), 0), cand], dim=1)
        ok = _first_occurrence(stream)                                   # padding entries are distinct negative sentinels
        ok[:, kept.shape[1]:] &= ~index.contains(hr[:, None], rr[:, None], cand)
        ok &= stream >= 0
        rank = torch.cumsum(ok, dim=1)
        take = ok & (rank <= k)
        n_ok = rank[:, -1].clamp(max=k)
        # compact the survivors to the left, stream order preserved
        pos = torch.where(take, rank - 1, torch.full_like(rank, k))
        buf = torch.full((rows.numel(), k + 1), -1, dtype=torch.int64, device=dev)
        buf.scatter_(1, pos, stream)
        done = n_ok >= k
        out[rows[done]] = buf[done, :k]
        if bool(done.all()):
            return out
        # rows still short keep their survivors (padded with -1) and draw again, with more candidates
        new_kept = torch.full((B, k), -1, dtype=torch.int64, device=dev)
        new_kept[rows[~done]] = buf[~done, :k]
        sentinel = -(1 + torch.arange(k, device=dev)).expand(B, k)       # distinct per column: never equal to a candidate or each other
        kept = torch.where(new_kept < 0, sentinel, new_kept)
        rows = rows[~done]
        C = min(4 * C, max(4 * k, 4096))
    raise RuntimeError("negative sampling did not converge")           # unreachable for k <= number of valid entities
