"""Copy-engine exchange over NVLink peer memory for the destination-sharded encoder (one process per GPU, one node).

Why not only NCCL: every compute kernel of the hot path is persistent with one CTA per SM (~200 KB of shared memory each), and
NCCL's kernels need SMs for the whole duration of a transfer.  Measured at 8 GPUs (profiles/multi_gpu_overlap_r02.md): the
exchanges cost 10.3 ms per step un-overlapped; issuing them asynchronously hid 2.6 ms; reserving SMs for NCCL or giving it fewer
CTAs only moved the cost between "exposed communication" and "slower compute".  A transfer that uses NO SM is a
``cudaMemcpyAsync`` between peer-mapped buffers: it runs on the copy engines.  So here

  * the all-gather of the layer input is 'every rank PULLS the other ranks' rows' from buffers the owners expose through CUDA IPC;
  * the reduce-scatter of the input-gradient partials is 'every rank pulls ITS slice of every other rank's partial table' into a
    staging buffer, followed by one local sum kernel (a reduction cannot be done by a copy engine);
  * ordering across processes uses no SM either: a rank announces "my producer kernels up to here are done" by copying an epoch
    number into a flag word of every peer (4-byte copy-engine writes, stream-ordered behind the producer) and waits for its peers
    with ``cuStreamWaitValue32`` on its own flag words.  (A tiny NCCL all-reduce would do, but its kernel cannot start while
    every SM is held by a persistent compute CTA -- exactly the stall this module exists to remove.)

Buffers are plain ``cudaMalloc`` allocations of libjmac_b200 (one allocation = one IPC handle), double-buffered per use, and
wrapped as torch tensors through ``__cuda_array_interface__`` so that the GEMM / edge kernels write straight into them.
Why two slots are enough: a rank reuses slot s two exchanges later; in between it has passed the handshake of the next exchange,
which a peer only joins after its own pulls of the previous one were consumed -- so nobody is still reading slot s.
Setup is collective and all-or-nothing: if any rank cannot allocate or map, every rank falls back to the NCCL path of dist.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class _Raw:
    """Zero-copy view of a device pointer for ``torch.as_tensor``."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}


class PeerBuf:
    """One buffer of ``nbytes`` on every rank; ``ptrs[g]`` addresses rank g's copy from THIS process (local pointer for g == rank)."""

    def __init__(self, nbytes: int, group, device):
        self.lib = _lib.load()
        self.group, self.device, self.nbytes = group, device, int(nbytes)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.local = ctypes.c_void_p()
        self.ptrs: List[Optional[int]] = [None] * self.world
        self._opened: List[int] = []
        with torch.cuda.device(device):
            check(self.lib.jmac_peer_alloc(self.nbytes, ctypes.byref(self.local)), "jmac_peer_alloc")
            handle = ctypes.create_string_buffer(64)
            check(self.lib.jmac_ipc_get_handle(self.local, handle), "jmac_ipc_get_handle")
        self.handle = handle.raw
        self.ptrs[self.rank] = int(self.local.value)

    def connect(self, handles: List[bytes]):
        with torch.cuda.device(self.device):
            for g, h in enumerate(handles):
                if g == self.rank:
                    continue
                p = ctypes.c_void_p()
                check(self.lib.jmac_ipc_open(h, ctypes.byref(p)), "jmac_ipc_open")
                self.ptrs[g] = int(p.value)
                self._opened.append(int(p.value))

    def tensor(self, shape, dtype=torch.float32) -> torch.Tensor:
        assert dtype == torch.float32
        t = torch.as_tensor(_Raw(self.ptrs[self.rank], shape), device=self.device)
        t._jmac_owner = self                     # the view must not outlive the allocation
        return t

    def close_imports(self):
        """Unmap the other ranks' copies (an exporter must not free its allocation while an importer still has it mapped)."""
        try:
            with torch.cuda.device(self.device):
                for p in self._opened:
                    self.lib.jmac_ipc_close(ctypes.c_void_p(p))
                self._opened = []
        except Exception:
            pass

    def free_local(self):
        try:
            with torch.cuda.device(self.device):
                if self.local:
                    self.lib.jmac_peer_free(self.local)
                    self.local = ctypes.c_void_p()
        except Exception:
            pass

    def close(self):
        self.close_imports()
        self.free_local()


class PeerExchange:
    """All-gather / reduce-scatter of ``[n_max, d]`` row blocks in the chunk-major gathered layout of ``dist.Shard``."""

    N_STREAMS = 4                                 # pulls are spread over a few streams so that several copy engines work at once

    def __init__(self, shard, d: int, group, device):
        self.shard, self.d, self.group, self.device = shard, int(d), group, device
        self.rank, self.world = shard.rank, shard.world
        self.lib = _lib.load()
        row = self.d * 4
        self.row_bytes = row
        ok = torch.ones(1, device=device)
        self.bufs: List[PeerBuf] = []
        self.sync_mode = os.environ.get("JMAC_PEER_SYNC", "flags")          # "nccl": a tiny all-reduce instead of the flag handshake (diagnostic)
        try:
            for nbytes in (shard.n_max * row, shard.n_max * row, shard.n_src * row, shard.n_src * row, 256):
                self.bufs.append(PeerBuf(nbytes, group, device))
            self.xsrc, self.part, self.flags = self.bufs[0:2], self.bufs[2:4], self.bufs[4]   # flags: world flag words + one scratch word
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) == 0.0:
            self.close()
            raise RuntimeError("peer buffers could not be allocated on every rank")
        handles = [None] * self.world
        dist.all_gather_object(handles, [b.handle for b in self.bufs], group=group)
        try:
            for i, b in enumerate(self.bufs):
                b.connect([h[i] for h in handles])
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) == 0.0:
            self.close()
            raise RuntimeError("peer buffers could not be mapped on every rank (no peer access?)")
        with torch.cuda.device(device):
            self.streams = [torch.cuda.Stream(device=device) for _ in range(self.N_STREAMS)]
            torch.as_tensor(_Raw(self.flags.ptrs[self.rank], (64,), "<i4"), device=device).zero_()
            self._epoch = 0
            self.x_t = [b.tensor((shard.n_max, self.d)) for b in self.xsrc]
            self.part_t = [b.tensor((shard.n_src, self.d)) for b in self.part]
            for t in self.x_t:
                t.zero_()                        # the padding rows behind n_loc stay zero for good
            self.stage = torch.empty(shard.chunks, self.world, shard.n_c, self.d, dtype=torch.float32, device=device)
            torch.cuda.synchronize()
        self._x_slot = 0
        self._p_slot = 0
        self._nccl_flag = torch.zeros(1, device=device)
        dist.barrier(group=group)
        # one handshake as a self-test (stream memory operations available, flag words reachable from every peer)
        try:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self._barrier_then(ev)
            for s in self.streams:
                s.synchronize()
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) == 0.0:
            self.close()
            raise RuntimeError("peer handshake failed: " + self.lib.jmac_last_error().decode())

    # ---- plumbing ---------------------------------------------------------------------------------------------------------
    def _barrier_then(self, after_event: torch.cuda.Event):
        """On stream 0: wait for ``after_event`` (this rank's producer), publish the next epoch to every peer, wait until every
        peer has published it: from then on every rank's producer has completed.  The other copy streams wait for stream 0."""
        s0 = self.streams[0]
        s0.wait_event(after_event)
        if self.sync_mode == "nccl":
            with torch.cuda.stream(s0):
                dist.all_reduce(self._nccl_flag, group=self.group)
            ev = torch.cuda.Event()
            ev.record(s0)
            for s in self.streams[1:]:
                s.wait_event(ev)
            return
        self._epoch = (self._epoch + 1) & 0xFFFFFFFF
        st = ctypes.c_void_p(s0.cuda_stream)
        lib = self.lib
        mine = self.flags.ptrs[self.rank]
        word = mine + 4 * self.world                                  # scratch word behind the flag words
        with torch.cuda.device(self.device):
            check(lib.jmac_stream_write_u32(ctypes.c_void_p(word), self._epoch, st), "jmac_stream_write_u32")
            for i in range(1, self.world):
                g = (self.rank + i) % self.world
                check(lib.jmac_memcpy_async(ctypes.c_void_p(self.flags.ptrs[g] + 4 * self.rank), ctypes.c_void_p(word), 4, st), "jmac_memcpy_async")
            for i in range(1, self.world):
                g = (self.rank + i) % self.world
                check(lib.jmac_stream_wait_geq_u32(ctypes.c_void_p(mine + 4 * g), self._epoch, st), "jmac_stream_wait_geq_u32")
        ev = torch.cuda.Event()
        ev.record(s0)
        for s in self.streams[1:]:
            s.wait_event(ev)

    def _copy(self, k: int, dst_ptr: int, src_ptr: int, nbytes: int):
        s = self.streams[k % self.N_STREAMS]
        with torch.cuda.device(self.device):
            check(self.lib.jmac_memcpy_async(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), nbytes, ctypes.c_void_p(s.cuda_stream)), "jmac_memcpy_async")

    def _done_events(self):
        evs = []
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            evs.append(ev)
        return evs

    # ---- all-gather ---------------------------------------------------------------------------------------------------------
    def all_gather(self, x_loc: torch.Tensor, out: torch.Tensor):
        """x_loc [n_loc, d] (current stream) -> out [world * n_max, d] (local tensor, gathered layout).  Returns, per chunk, the
        events the consumer stream has to wait for."""
        sh = self.shard
        slot = self._x_slot
        self._x_slot ^= 1
        main = torch.cuda.current_stream(self.device)
        self.x_t[slot][: sh.n_loc].copy_(x_loc)
        ready = torch.cuda.Event()
        ready.record(main)
        self._barrier_then(ready)
        chunk_bytes = sh.n_c * self.row_bytes
        out_ptr = out.data_ptr()
        per_chunk = []
        for c in range(sh.chunks):
            r0, _ = sh.chunk_rows(c)
            for i in range(self.world):
                g = (self.rank + i) % self.world                                  # start with the own rows, then ring order: spreads the load over the links
                src = self.xsrc[slot].ptrs[g] + c * chunk_bytes
                dst = out_ptr + (r0 + g * sh.n_c) * self.row_bytes
                self._copy(i, dst, src, chunk_bytes)
            per_chunk.append(self._done_events())
        return per_chunk

    # ---- reduce-scatter -----------------------------------------------------------------------------------------------------
    def partial_table(self) -> torch.Tensor:
        """The [world * n_max, d] table the source pass / chunk GEMMs of this backward write the input-gradient partials into."""
        self._cur_part = self._p_slot
        self._p_slot ^= 1
        self._pulls = []
        return self.part_t[self._cur_part]

    def scatter_chunk(self, c: int):
        """Chunk c of the partial table is complete on the current stream: pull this rank's slice of it from every rank."""
        sh = self.shard
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self._barrier_then(ready)
        chunk_bytes = sh.n_c * self.row_bytes
        r0, _ = sh.chunk_rows(c)
        stage_ptr = self.stage.data_ptr()
        for i in range(self.world):
            g = (self.rank + i) % self.world
            src = self.part[self._cur_part].ptrs[g] + (r0 + self.rank * sh.n_c) * self.row_bytes
            dst = stage_ptr + ((c * self.world + g) * sh.n_c) * self.row_bytes
            self._copy(i, dst, src, chunk_bytes)
        self._pulls.append(self._done_events())

    def wait_pulls(self):
        main = torch.cuda.current_stream(self.device)
        for evs in self._pulls:
            for ev in evs:
                main.wait_event(ev)

    def scatter_finish(self, acc: torch.Tensor):
        """acc [n_loc, d] += sum over the ranks of this rank's rows of the partial tables (ranks added in ascending order:
        deterministic).  The current stream waits for the pulls here, not earlier."""
        sh = self.shard
        main = torch.cuda.current_stream(self.device)
        for evs in self._pulls:
            for ev in evs:
                main.wait_event(ev)
        self._pulls = []
        assert acc.is_contiguous() and acc.dtype == torch.float32 and acc.shape == (sh.n_loc, self.d)
        st = ctypes.c_void_p(main.cuda_stream)
        per = sh.n_c * self.d
        with torch.cuda.device(self.device):
            for c in range(sh.chunks):
                rows = min(sh.n_c, sh.n_loc - c * sh.n_c)
                if rows <= 0:
                    break
                check(self.lib.jmac_peer_sum(ctypes.c_void_p(acc.data_ptr() + c * per * 4), ctypes.c_void_p(self.stage.data_ptr() + c * self.world * per * 4),
                                             self.world, per, rows * self.d, 1, st), "jmac_peer_sum")
        return acc

    def close(self, collective: bool = False):
        """Release the buffers.  ``collective``: every rank of the group calls this at the same point; the mappings of the other ranks'
        buffers are closed everywhere (barrier) before any rank frees its own allocations."""
        bufs, self.bufs = self.bufs, []
        for b in bufs:
            b.close_imports()
        if collective and bufs:
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        for b in bufs:
            b.free_local()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
