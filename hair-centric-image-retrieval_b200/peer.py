"""Exchange of per-rank result blocks over NVLink peer memory (csrc/peer.cu): the path's own
all-gather.  Every rank stores its block straight into a slot of EVERY rank's region with plain
stores (the regions are cudaMalloc'd by the library, exported with CUDA IPC and mapped by all ranks
of the box) and bumps an arrival counter behind a system-scope fence; the consumer kernel on each
rank spins until all blocks of the step have landed.  In the search step the producer is the TAIL
OF K3 (every query's CTA stores its own results into the peers: ``fill_tail``) and the consumer is
the fused wait + merge + vote kernel (``hcir_peer_merge_vote``) or the one-warp wait kernel
(``enqueue_wait``); ``exchange`` is the standalone push + wait for blocks that already sit in
memory.  No collective-library call sits on the data path, and because the completed-step counter
lives in device memory the exchange is captured into the step's CUDA graph and replayed unchanged.

``torch.distributed`` is used once, at construction, to trade the 64-byte IPC handles (SURVEY.md
section 8e "later fusion"; no reference analogue -- the reference's kNN is single-process CPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

_HDR_WORDS = 64          # 512-byte header, int64 words (peer.cu)
_HDR_META, _HDR_STEP, _HDR_ERR, _PEER_MAX = 16, 48, 49, 16


class _DeviceBytes:
    """A raw device allocation presented through the CUDA array interface (torch.as_tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerUnavailable(RuntimeError):
    """Raised on EVERY rank of the group when any rank could not set the peer regions up."""


def _all_ok(ok: bool, group, device) -> bool:
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def probe(group=None, device=None) -> bool:
    """Collective self-test: one tiny channel, one exchanged step, contents checked on every rank.
    True on all ranks or False on all ranks."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    try:
        xc = PeerExchange(1024, group=group, device=device, timeout_s=2.0)
    except PeerUnavailable:
        return False
    ok = True
    try:
        with torch.cuda.device(device):
            blk = torch.full((1024,), xc.rank + 1, dtype=torch.uint8, device=device)
            meta = torch.tensor([100 + xc.rank], dtype=torch.int32, device=device)
            for _ in range(3):   # both parities
                xc.exchange(blk, meta)
                ok = ok and xc.metas() == [100 + r for r in range(xc.world)]
                g = xc.gathered()[:, :1024]
                want = torch.arange(1, xc.world + 1, dtype=torch.uint8, device=device)[:, None].expand(-1, 1024)
                ok = ok and bool(torch.equal(g, want))
    except RuntimeError:
        ok = False
    ok = _all_ok(ok, group, device)
    xc.close()
    return ok


class PeerExchange:
    """One exchange channel: ``slot_bytes`` per rank and step.  Construction is COLLECTIVE over
    ``group`` (IPC handle all-gather + barrier) and must happen outside CUDA-graph capture."""

    def __init__(self, slot_bytes: int, *, group=None, device=None, timeout_s: float | None = None):
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _PEER_MAX:
            raise ValueError(f"peer exchange supports up to {_PEER_MAX} ranks, got {self.world}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slot_bytes = int(-(-int(slot_bytes) // 16) * 16)
        self.stride = -(-self.slot_bytes // 256) * 256
        # how long a wait kernel spins for a late peer before it reports instead of hanging the GPU
        # (ranks of one job reach the same step seconds apart at most; HCIR_PEER_TIMEOUT_S overrides)
        # 120 s by default: long enough for a peer that is finishing a batch of uncertified queries on
        # the exact fp32 kernel or sitting in a debugger breakpoint-free stall, short enough not to
        # look like a hung GPU.  After a timeout the channel is dead (the error word is sticky): the
        # owner must close() it and build a new one -- ShardedGallery / QueryShardedGallery.close().
        if timeout_s is None:
            timeout_s = float(os.environ.get("HCIR_PEER_TIMEOUT_S", "120"))
        self.timeout_ns = int(timeout_s * 1e9)
        total = int(self.lib.hcir_peer_region_bytes(self.world, self.slot_bytes))
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            # every collective below is reached by every rank whatever failed locally, and a failure
            # anywhere raises PeerUnavailable everywhere
            why = None
            rc = self.lib.hcir_peer_alloc(total, C.byref(ptr), handle)
            if rc != 0:
                why = f"peer_alloc: {_lib.last_error()}"
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw if rc == 0 else None, group=group)
            self._local_ptr = int(ptr.value) if rc == 0 else 0
            ptrs = []
            if all(h is not None for h in handles):
                for r in range(self.world):
                    if r == self.rank:
                        ptrs.append(self._local_ptr)
                        continue
                    p = C.c_void_p()
                    if self.lib.hcir_peer_open(C.create_string_buffer(handles[r], 64), C.byref(p)) != 0:
                        why = f"peer_open(rank {r}): {_lib.last_error()}"
                        break
                    ptrs.append(int(p.value))
            else:
                why = why or "a peer could not allocate / export its region"
            if not _all_ok(why is None, group, self.device):
                for r, p in enumerate(ptrs):
                    if r != self.rank:
                        self.lib.hcir_peer_close(p)
                dist.barrier(group)
                if self._local_ptr:
                    self.lib.hcir_peer_free(self._local_ptr)
                raise PeerUnavailable(f"rank {self.rank}: {why or 'a peer failed to map the regions'}")
            self._ptrs = ptrs
            self._regions = (C.c_void_p * self.world)(*ptrs)
            self._mem = _DeviceBytes(self._local_ptr, total)
            self.local = torch.as_tensor(self._mem, device=self.device)        # uint8 view of the local region
            self.step = torch.zeros((), dtype=torch.int64, device=self.device)  # device-side step counter
            self._hdr_host = torch.empty((_HDR_WORDS,), dtype=torch.int64, pin_memory=True)
            torch.cuda.synchronize(self.device)
        dist.barrier(group)   # every region is zeroed and mapped everywhere before the first push
        self.host_step = 0
        self.block_bytes = None
        self._closed = False

    # ------------------------------------------------------------------ data path (capturable)
    def _check_block(self, nbytes: int):
        if nbytes % 8 or nbytes > self.slot_bytes:
            raise ValueError(f"peer exchange block must be a multiple of 8 bytes and <= {self.slot_bytes}")
        if self.block_bytes not in (None, nbytes):   # one block shape per channel
            raise ValueError(f"peer exchange channel carries blocks of {self.block_bytes} bytes, got {nbytes}")
        self.block_bytes = nbytes

    def fill_tail(self, tail, payload: int, nbytes: int):
        """Make K3 the producer of this channel: the peer fields of an hcir_tail_t (_lib.Tail)."""
        self._check_block(nbytes)
        tail.world, tail.rank, tail.payload = self.world, self.rank, int(payload)
        tail.slot_bytes = self.slot_bytes
        for r, p in enumerate(self._ptrs):
            tail.regions[r] = p
        tail.step = self.step.data_ptr()

    def enqueue_wait(self):
        """Consumer, wait only: every rank's block of the step has landed; completes the step."""
        _lib.check(self.lib.hcir_peer_wait(self._local_ptr, self.world, self.step.data_ptr(), self.timeout_ns,
                                           torch.cuda.current_stream().cuda_stream), "peer_wait")
        if not torch.cuda.is_current_stream_capturing():
            self.host_step += 1

    def exchange(self, block: torch.Tensor, meta: torch.Tensor | None = None):
        """Standalone producer + consumer: push ``block`` (contiguous bytes, multiple of 16) + the
        int32 ``meta`` word into every rank's region, wait for every rank's block of this step."""
        nbytes = block.numel() * block.element_size()
        if not block.is_contiguous() or nbytes % 16:
            raise ValueError("peer exchange block must be contiguous and a multiple of 16 bytes")
        self._check_block(nbytes)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.hcir_peer_push(block.data_ptr(), nbytes, self._regions, self.world, self.rank,
                                           self.slot_bytes, self.step.data_ptr(),
                                           meta.data_ptr() if meta is not None else None, st), "peer_push")
        self.enqueue_wait()

    def note_replay(self):
        """A CUDA graph holding one captured step of this channel was replayed."""
        self.host_step += 1

    # ------------------------------------------------------------------ reading the result
    @property
    def local_ptr(self) -> int:
        return self._local_ptr

    def header(self) -> np.ndarray:
        """Synchronising read of the local header; raises if a wait timed out or the device step
        differs from the host's count (a rank skipped or repeated a step)."""
        self._hdr_host.copy_(self.local[: _HDR_WORDS * 8].view(torch.int64), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = self._hdr_host.numpy()
        if h[_HDR_ERR] != 0:
            raise RuntimeError(f"peer exchange: rank {self.rank} timed out waiting for a peer's block at step "
                               f"{int(h[_HDR_ERR])}")
        if int(h[_HDR_STEP]) != self.host_step:
            raise RuntimeError(f"peer exchange: device step {int(h[_HDR_STEP])} != host step {self.host_step}")
        return h

    def header_async(self):
        """Enqueue the header read of the step just issued into a pinned slot (ring of 8); returns
        (slot, step) for ``metas_of`` once the caller has synchronised with the stream."""
        ring = self.__dict__.get("_hdr_ring")
        if ring is None:
            ring = self._hdr_ring = torch.zeros((8, _HDR_WORDS), dtype=torch.int64, pin_memory=True)
        slot = ring[self.host_step % 8]
        slot.copy_(self.local[: _HDR_WORDS * 8].view(torch.int64), non_blocking=True)
        return slot, self.host_step

    def metas_of(self, slot: torch.Tensor, step: int) -> list[int]:
        h = slot.numpy()
        if h[_HDR_ERR] != 0:
            raise RuntimeError(f"peer exchange: rank {self.rank} timed out waiting for a peer's block at step "
                               f"{int(h[_HDR_ERR])}")
        if int(h[_HDR_STEP]) != step:
            raise RuntimeError(f"peer exchange: device step {int(h[_HDR_STEP])} != host step {step}")
        par = step & 1
        return [int(v) for v in h[_HDR_META + par * _PEER_MAX: _HDR_META + par * _PEER_MAX + self.world]]

    def metas(self, header: np.ndarray | None = None) -> list[int]:
        h = self.header() if header is None else header
        par = self.host_step & 1
        return [int(v) for v in h[_HDR_META + par * _PEER_MAX: _HDR_META + par * _PEER_MAX + self.world]]

    def gathered(self) -> torch.Tensor:
        """[world, stride] uint8 view of the blocks of the LAST completed step (valid until the
        step after next)."""
        off = int(self.lib.hcir_peer_slot_offset(self.world, self.slot_bytes, self.host_step & 1, 0))
        return self.local[off: off + self.world * self.stride].view(self.world, self.stride)

    def close(self):
        if self._closed:
            return
        self._closed = True
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r, p in enumerate(self._ptrs):
                if r != self.rank:
                    self.lib.hcir_peer_close(p)
            try:
                dist.barrier(self.group)   # nobody frees a region a peer still maps
            except Exception:
                pass
            self.local = None
            self.lib.hcir_peer_free(self._local_ptr)
