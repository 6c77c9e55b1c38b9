"""CUDA-graph capture of the reference's training steps at the reference's own sizes.

At DBP-5L scale (5-14 k entities, 13-73 k edges, B = 1000, K = 25; train.py:328-378) one
``completion_loss`` / ``alignment_loss`` step is ~150 kernel launches of a few microseconds each: the step is
bound by the host's launch path (Python -> ctypes / ATen -> driver), not by the GPU.  The graph is static
between two triple-transfer rounds (train.py:447-485 reuses one edge set for every step of an epoch), so the
whole step -- both encoders forward, scorer, backward, Adam -- is captured ONCE and replayed with new batch
indices copied into static buffers.  B200-first: CUDA streams and graphs instead of a tracing compiler.

The captured step is the same code path as the eager one (``JMAC.completion_loss`` -> libjmac_b200.so), so
results are bit-identical to calling the model directly with the same parameters.

Not a replacement for the drop-in API: ``model.completion_loss(...)`` keeps working un-captured.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import functional as JF
from .functional import as_index


def _device_feed(feeddict: Dict, device) -> Dict:
    """Index arrays of the feed dictionary as device tensors (a pageable host->device copy cannot be captured)."""
    out = dict(feeddict)
    for k in ("links", "neg_left", "neg_right", "neg2_left", "neg2_right"):
        v = out.get(k)
        if v is None:
            continue
        if torch.is_tensor(v):
            out[k] = v.to(device=device, dtype=torch.int64)
        else:
            a = np.asarray(v)
            out[k] = torch.from_numpy(np.ascontiguousarray(a.astype(np.int64))).to(device)
    return out


class _GraphedStep:
    """Common machinery: warm up on a side stream, capture forward + backward + optimizer.step(), restore the
    model / optimizer state the warm-up consumed (in place: the graph holds the tensors' addresses)."""

    def __init__(self, model, optimizer, warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA graphs need a CUDA device")
        if optimizer is not None and not optimizer.defaults.get("capturable", False):
            raise ValueError("the optimizer's step counter must live on the device: construct it with capturable=True "
                             "(torch.optim.Adam(params, lr=..., capturable=True))")
        self.model, self.optimizer = model, optimizer
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._warmup = int(warmup)

    def _loss(self) -> torch.Tensor:
        raise NotImplementedError

    def _one_step(self):
        if self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self._loss()
        if torch.is_tensor(loss) and loss.requires_grad:
            loss.backward()
            if self.optimizer is not None:
                self.optimizer.step()
        return loss

    def _capture(self):
        model, opt = self.model, self.optimizer
        params = [p for p in model.parameters()]
        buffers = [b for b in model.buffers()]
        saved_p = [p.detach().clone() for p in params]
        saved_b = [b.detach().clone() for b in buffers]
        had_state = opt is not None and len(opt.state) > 0
        saved_state = None
        if had_state:
            saved_state = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in opt.state.items()}
        self._plans = []                      # the graph plans the captured kernels read: kept alive with the graph
        JF._PLAN_PINS = self._plans
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self._warmup):
                    self._one_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if opt is not None:
                opt.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._one_step()
        finally:
            JF._PLAN_PINS = None
        # the warm-up steps and the capture must not be visible: put parameters, buffers and optimizer state back IN PLACE
        with torch.no_grad():
            for p, s in zip(params, saved_p):
                p.copy_(s)
            for b, s in zip(buffers, saved_b):
                b.copy_(s)
            if opt is not None:
                for p, st in opt.state.items():
                    old = saved_state.get(id(p)) if saved_state else None
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()
        torch.cuda.synchronize()

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_loss


class GraphedCompletionStep(_GraphedStep):
    """One ``train_completion_component`` iteration (train.py:341-358) as a CUDA graph:
    ``optimizer.zero_grad(); loss = model.completion_loss(data, ...); loss.backward(); optimizer.step()``.

    ``step(data)`` copies ``data['batch_h' | 'batch_r' | 'batch_t']`` (device tensors of ``batch_rows`` = (1+K)*B
    entries, the layout of train.py:347-352) into the captured buffers, replays, and returns the loss tensor
    (shape [1], still on the device: ``.item()`` it when the value is needed, as train.py:357 does).
    The edge tensors and the feed dictionary are fixed for the lifetime of the object -- build a new one after a
    triple-transfer round changes the graph (train.py:471-472)."""

    def __init__(self, model, optimizer, edge_index1, edge_type1, edge_index2, edge_type2, feeddict, source: bool, batch_rows: int,
                 warmup: int = 3):
        super().__init__(model, optimizer, warmup)
        dev = edge_index1.device
        self.edges = (edge_index1, edge_type1, edge_index2, edge_type2)
        self.feed = _device_feed(feeddict, dev)
        self.source = bool(source)
        self.rows = int(batch_rows)
        # valid placeholder indices (0) until the first real batch arrives
        self.data = {k: torch.zeros(self.rows, dtype=torch.int64, device=dev) for k in ("batch_h", "batch_r", "batch_t")}
        self._capture()

    def _loss(self):
        return self.model.completion_loss(self.data, *self.edges, self.feed, self.source)

    def __call__(self, data) -> torch.Tensor:
        for k, buf in self.data.items():
            src = as_index(data[k], buf.device).reshape(-1)
            if src.numel() != self.rows:
                raise ValueError(f"{k} has {src.numel()} entries, the captured step was built for {self.rows} "
                                 "(train.py:340-341 skips the ragged last batch for the same reason)")
            buf.copy_(src, non_blocking=True)
        return self.replay()


class GraphedAlignmentStep(_GraphedStep):
    """``train_alignment_component`` (train.py:367-378) as a CUDA graph.  The negatives of the feed dictionary are
    refreshed in place with ``update_negatives`` when seed enlargement recomputes them (same lengths)."""

    def __init__(self, model, optimizer, edge_index1, edge_type1, edge_index2, edge_type2, feeddict, warmup: int = 3):
        super().__init__(model, optimizer, warmup)
        self.edges = (edge_index1, edge_type1, edge_index2, edge_type2)
        self.feed = _device_feed(feeddict, edge_index1.device)
        self._capture()

    def _loss(self):
        return self.model.alignment_loss(self.feed, *self.edges)

    def update_negatives(self, **arrays):
        for k, v in arrays.items():
            src = as_index(v, self.feed[k].device).reshape(-1)
            if src.numel() != self.feed[k].numel():
                raise ValueError(f"{k}: length changed, rebuild the captured step")
            self.feed[k].copy_(src.view_as(self.feed[k]))

    def __call__(self) -> torch.Tensor:
        return self.replay()
