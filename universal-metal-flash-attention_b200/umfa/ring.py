"""Context-parallel causal attention over the GPUs of one node (ring attention, zig-zag partition).

The reference is a single-device library ("Reference support: none"); this is the additive multi-GPU layer
BASELINE.json's north_star asks for: one process per GPU, the sequence split into 2G chunks, rank r owning chunks r and
2G-1-r so causal work is balanced; Q stays put while the K/V chunks travel round the ring (G-1 hops of
torch.distributed send/recv -- NCCL over NVLink on GPUs, gloo in the CPU tests) overlapped with the attention of the
previous block.  Every (q chunk, kv chunk) pair is either fully visible, the causal diagonal, or fully hidden and
skipped before any work is issued, so the kernels only ever see the two mask modes they already have.  Partial results
carry L = log2-domain logsumexp (the library's A3 convention) and are merged as
    L = log2(2^L1 + 2^L2),  O = O1 2^(L1-L) + O2 2^(L2-L).

The driver is generic over a small backend (attend / merge / exchange) so the schedule and merge logic run unchanged
on CPU tensors in the world_size-2 gloo tests; `CudaBackend` binds it to libMFAFFI.so.
"""
import ctypes
from typing import List, Tuple

import numpy as np

LOG2E = 1.4426950408889634


# ------------------------------------------------------------------------------------------------ partition / schedule
def chunk_ids(rank: int, world: int) -> Tuple[int, int]:
    """Global chunk indices (of 2*world equal chunks) owned by `rank`: (low, high)."""
    return rank, 2 * world - 1 - rank


def shard_sequence(x, rank: int, world: int, dim: int = 2):
    """Slices [.., N, ..] into this rank's (low, high) chunks along `dim`; N must divide by 2*world."""
    n = x.shape[dim]
    if n % (2 * world):
        raise ValueError(f"sequence length {n} must be a multiple of 2*world={2 * world}")
    c = n // (2 * world)
    lo, hi = chunk_ids(rank, world)
    idx = [slice(None)] * x.ndim
    idx[dim] = slice(lo * c, (lo + 1) * c)
    a = x[tuple(idx)]
    idx[dim] = slice(hi * c, (hi + 1) * c)
    b = x[tuple(idx)]
    return a, b


def step_schedule(rank: int, world: int, step: int) -> List[Tuple[int, int, bool]]:
    """Work of `rank` at ring step `step` (the K/V on hand come from rank (rank - step) mod world).

    Returns [(q_chunk, kv_chunk, causal)] with chunk = 0 (low) / 1 (high) of the local / visiting pair; pairs whose
    keys all lie in the future of the queries are omitted.  Every step costs two chunk-pairs of work on every rank."""
    src = (rank - step) % world
    q_lo, q_hi = chunk_ids(rank, world)
    k_lo, k_hi = chunk_ids(src, world)
    out = []
    for qi, qc in ((0, q_lo), (1, q_hi)):
        for ki, kc in ((0, k_lo), (1, k_hi)):
            if kc < qc:
                out.append((qi, ki, False))
            elif kc == qc:
                out.append((qi, ki, True))
    return out


def visible_pairs_causal(n: int) -> int:
    return n * (n + 1) // 2


# ------------------------------------------------------------------------------------------------ generic driver
def ring_attention_forward(backend, q_pair, k_pair, v_pair, rank: int, world: int, scale: float):
    """q_pair/k_pair/v_pair: (low, high) chunk tensors [B, H, C, D] of this rank.  Returns ((o_lo, l_lo), (o_hi, l_hi)):
    fp32 O [B,H,C,D] and L [B,H,C] (log2 units) of the two local query chunks."""
    acc = [None, None]
    kv_cur = backend.pack_kv(k_pair, v_pair)
    for step in range(world):
        handle = None
        if step + 1 < world:
            handle = backend.exchange_start(kv_cur, (rank + 1) % world, (rank - 1) % world, step)
        k_c, v_c = backend.unpack_kv(kv_cur)
        for qi, ki, causal in step_schedule(rank, world, step):
            if acc[qi] is None:
                acc[qi] = backend.attend(q_pair[qi], k_c[ki], v_c[ki], causal, scale, out=None)
            else:
                part = backend.attend(q_pair[qi], k_c[ki], v_c[ki], causal, scale, out="scratch")
                backend.merge(acc[qi], part)
        if handle is not None:
            kv_cur = backend.exchange_finish(handle, step)
    backend.finish()
    return acc[0], acc[1]


def merge_partials_numpy(o_acc, l_acc, o_part, l_part):
    """Reference merge (numpy, in place) of the rule in the module docstring; -inf rows contribute nothing."""
    m = np.maximum(l_acc, l_part)
    safe = np.where(np.isfinite(m), m, 0.0)
    wa = np.where(np.isfinite(l_acc), np.exp2(l_acc - safe), 0.0)
    wb = np.where(np.isfinite(l_part), np.exp2(l_part - safe), 0.0)
    tot = wa + wb
    inv = np.where(tot > 0, 1.0 / np.where(tot > 0, tot, 1.0), 0.0)
    o_acc[...] = (o_acc * (wa * inv)[..., None] + o_part * (wb * inv)[..., None]).astype(o_acc.dtype)
    l_acc[...] = np.where(tot > 0, safe + np.log2(np.where(tot > 0, tot, 1.0)), -np.inf).astype(l_acc.dtype)


# ------------------------------------------------------------------------------------------------ CPU backend (tests)
class HostBackend:
    """numpy tensors + torch.distributed (gloo) point-to-point; `attend_fn(q, k, v, causal, scale) -> (o, lse)` is
    injected by the caller (the tests pass the CPU oracle).  Exists to exercise the schedule, the exchange ordering and
    the merge on a CPU-only machine."""

    def __init__(self, attend_fn, dist=None):
        self.attend_fn = attend_fn
        self.dist = dist

    def pack_kv(self, k_pair, v_pair):
        return np.ascontiguousarray(np.stack([k_pair[0], k_pair[1], v_pair[0], v_pair[1]]))

    def unpack_kv(self, buf):
        return (buf[0], buf[1]), (buf[2], buf[3])

    def attend(self, q, k, v, causal, scale, out=None):
        o, l = self.attend_fn(q, k, v, causal, scale)
        return [np.array(o, np.float32), np.array(l, np.float32)]

    def merge(self, acc, part):
        merge_partials_numpy(acc[0], acc[1], part[0], part[1])

    def exchange_start(self, buf, dst, src, step):
        import torch
        send = torch.from_numpy(buf)
        recv = torch.empty_like(send)
        reqs = [self.dist.isend(send, dst), self.dist.irecv(recv, src)]
        return reqs, recv, send

    def exchange_finish(self, handle, step):
        reqs, recv, _send = handle
        for r in reqs:
            r.wait()
        return recv.numpy()

    def finish(self):
        pass


# ------------------------------------------------------------------------------------------------ CUDA backend
class CudaBackend:
    """torch CUDA tensors for storage and NCCL plumbing; attention and merge go through libMFAFFI.so on the compute
    stream (mfa_attention_forward_ex / mfa_merge_partials with device handles), K/V hops through
    torch.distributed.batch_isend_irecv on a side stream with two K/V buffers so hop s+1 overlaps the attention of hop s."""

    def __init__(self, ctx, dist, device, dtype="bf16"):
        import torch
        from . import _ffi
        self.torch, self.ctx, self.dist, self.device = torch, ctx, dist, device
        self.lib = _ffi._lib
        self.prec = {"bf16": 1, "fp16": 0}[dtype]
        self.compute = torch.cuda.current_stream(device)
        self.comm = torch.cuda.Stream(device)
        self.stream_ptr = ctypes.c_void_p(self.compute.cuda_stream)
        self.kv_bufs = [None, None]
        self.compute_done = [None, None]
        self.scratch = {}
        self.launches = 0

    def _buf(self, t):
        from .core import MFABuffer
        return MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size())

    def pack_kv(self, k_pair, v_pair):
        torch = self.torch
        buf = torch.stack([k_pair[0], k_pair[1], v_pair[0], v_pair[1]]).contiguous()
        self.kv_bufs[0] = buf
        self.kv_bufs[1] = torch.empty_like(buf)
        self.cur = 0
        return buf

    def unpack_kv(self, buf):
        return (buf[0], buf[1]), (buf[2], buf[3])

    def attend(self, q, k, v, causal, scale, out=None):
        torch = self.torch
        B, H, C, D = q.shape
        Skv = k.shape[2]
        if out == "scratch":
            key = (B, H, C, D)
            if key not in self.scratch:
                self.scratch[key] = [torch.empty(B, H, C, D, device=self.device, dtype=torch.float32),
                                     torch.empty(B, H, C, device=self.device, dtype=torch.float32)]
            o, l = self.scratch[key]
        else:
            o = torch.empty(B, H, C, D, device=self.device, dtype=torch.float32)
            l = torch.empty(B, H, C, device=self.device, dtype=torch.float32)
        bufs = [self._buf(t) for t in (q, k, v, o, l)]
        rc = self.lib.mfa_attention_forward_ex(self.ctx.handle, *[b.handle for b in bufs], B, C, Skv, H, D, scale, causal,
                                               -1, self.prec, 2, None, 0, None, None, 0, 0, 0, self.stream_ptr)
        for b in bufs:
            b.close()
        if rc != 0:
            raise RuntimeError(f"mfa_attention_forward_ex failed: {rc}")
        self.launches += 1
        return [o, l]

    def merge(self, acc, part):
        rows = acc[1].numel()
        D = acc[0].shape[-1]
        bufs = [self._buf(t) for t in (acc[0], acc[1], part[0], part[1])]
        rc = self.lib.mfa_merge_partials(self.ctx.handle, *[b.handle for b in bufs], rows, D, self.stream_ptr)
        for b in bufs:
            b.close()
        if rc != 0:
            raise RuntimeError(f"mfa_merge_partials failed: {rc}")
        self.launches += 1

    def exchange_start(self, buf, dst, src, step):
        torch, dist = self.torch, self.dist
        nxt = self.kv_bufs[self.cur ^ 1]
        # the receive buffer was last read by the attention of the previous step
        if self.compute_done[self.cur ^ 1] is not None:
            self.comm.wait_event(self.compute_done[self.cur ^ 1])
        ready = torch.cuda.Event()
        ready.record(self.compute)              # buf was produced (or received and waited for) on the compute stream
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            ops = [dist.P2POp(dist.isend, buf, dst), dist.P2POp(dist.irecv, nxt, src)]
            reqs = dist.batch_isend_irecv(ops)
        return reqs, nxt

    def exchange_finish(self, handle, step):
        torch = self.torch
        reqs, nxt = handle
        done = torch.cuda.Event()
        done.record(self.compute)               # attention of this step (reads kv_bufs[cur]) enqueued up to here
        self.compute_done[self.cur] = done
        with torch.cuda.stream(self.comm):
            for r in reqs:
                r.wait()
            arrived = torch.cuda.Event()
            arrived.record(self.comm)
        self.compute.wait_event(arrived)
        self.cur ^= 1
        return nxt

    def finish(self):
        pass
