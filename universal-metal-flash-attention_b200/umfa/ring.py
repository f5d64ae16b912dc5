"""Context-parallel causal attention over the GPUs of one node (ring attention, zig-zag partition).

The reference is a single-device library ("Reference support: none"); this is the additive multi-GPU layer
BASELINE.json's north_star asks for: one process per GPU, the sequence split into 2G chunks, rank r owning chunks r and
2G-1-r so causal work is balanced; Q stays put while the K/V chunks travel round the ring (G-1 hops of
torch.distributed send/recv -- NCCL over NVLink on GPUs, gloo in the CPU tests) overlapped with the attention of the
previous block.  Every (q chunk, kv chunk) pair is either fully visible, the causal diagonal, or fully hidden; with the rank's two
chunks stored next to each other the visible pairs of a step form ONE rectangular problem (step_plan), so a ring step is
a single kernel launch.  Partial results carry L = log2-domain logsumexp (the library's A3 convention) and are merged as
    L = log2(2^L1 + 2^L2),  O = O1 2^(L1-L) + O2 2^(L2-L)
inside the attention kernel's epilogue (mfa_attention_forward_accumulate): no partial O is written to HBM.

The driver is generic over a small backend (attend / attend_accumulate / exchange) so the schedule and merge logic run unchanged
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
    """Work of `rank` at ring step `step` (the K/V on hand come from rank (rank - step) mod world), chunk by chunk.

    Returns [(q_chunk, kv_chunk, causal)] with chunk = 0 (low) / 1 (high) of the local / visiting pair; pairs whose
    keys all lie in the future of the queries are omitted.  Every step costs two chunk-pairs of work on every rank.
    (The definition the fused plan below is tested against.)"""
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


def step_plan(rank: int, world: int, step: int) -> Tuple[Tuple[int, int], Tuple[int, int], bool]:
    """The same work as `step_schedule`, as ONE rectangular attention problem over the rank's sequence-adjacent
    [low | high] layout: ((q_row0, q_rows), (kv_row0, kv_rows), causal) in units of chunks.

      step 0 (own K/V): the zig-zag pair is causal on local indices (low chunk first)      -> 2 x 2 chunks, causal
      source rank < rank: both local query chunks see the visitor's low chunk only          -> 2 x 1 chunks, full
      source rank > rank: only the high query chunk sees the visitor, and sees both chunks  -> 1 x 2 chunks, full
    One launch per step instead of two or three, and no tile is launched only to be skipped."""
    src = (rank - step) % world
    if src == rank:
        return (0, 2), (0, 2), True
    if src < rank:
        return (0, 2), (0, 1), False
    return (1, 1), (0, 2), False


def visible_pairs_causal(n: int) -> int:
    return n * (n + 1) // 2


# ------------------------------------------------------------------------------------------------ generic driver
def ring_attention_forward(backend, q_pair, k_pair, v_pair, rank: int, world: int, scale: float):
    """q_pair/k_pair/v_pair: (low, high) chunk tensors [B, H, C, D] of this rank.  Returns ((o_lo, l_lo), (o_hi, l_hi)):
    fp32 O [B,H,C,D] and L [B,H,C] (log2 units) of the two local query chunks (views of one [B,H,2C,.] result)."""
    C = q_pair[0].shape[2]
    q_all = backend.cat_seq(q_pair)                      # [B, H, 2C, D], low chunk first
    kv_cur = backend.pack_kv(k_pair, v_pair)             # [2, B, H, 2C, D]: K and V, low | high adjacent along the sequence
    acc = None
    for step in range(world):
        handle = None
        if step + 1 < world:
            handle = backend.exchange_start(kv_cur, (rank + 1) % world, (rank - 1) % world, step)
        k_all, v_all = backend.unpack_kv(kv_cur)
        (q0, qn), (k0, kn), causal = step_plan(rank, world, step)
        q = q_all[:, :, q0 * C:(q0 + qn) * C]
        k = k_all[:, :, k0 * C:(k0 + kn) * C]
        v = v_all[:, :, k0 * C:(k0 + kn) * C]
        if acc is None:
            acc = backend.attend(q, k, v, causal, scale)             # step 0 covers every local row
        else:
            backend.attend_accumulate(acc, q0 * C, q, k, v, causal, scale)
        if handle is not None:
            kv_cur = backend.exchange_finish(handle, step)
    backend.finish()
    o, l = acc
    return (o[:, :, :C], l[:, :, :C]), (o[:, :, C:], l[:, :, C:])


def ring_attention_backward(backend, q_pair, k_pair, v_pair, o_pair, l_pair, do_pair, rank: int, world: int, scale: float):
    """Backward of `ring_attention_forward` ("ring backward").  Inputs are this rank's (low, high) chunks of Q, K, V,
    of the forward's O (fp32) and L (log2 units) and of the upstream gradient dO; returns ((dq_lo, dq_hi), (dk_lo, dk_hi),
    (dv_lo, dv_hi)), fp32.

    With the FINAL L and D = scale * rowsum(dO * O) of a query row, the flash backward of a (query block, key block) rectangle
    is an exact partial sum (P = exp2(S c - L), dS = P (dP - D)): the same rectangles as the forward (`step_plan`) are visited,
    dQ accumulates locally, and dK / dV accumulators TRAVEL WITH their K / V round the ring -- one more hop than the forward
    brings them home.  One backward launch pair (dK/dV + dQ kernel) per step through `backend.backward`."""
    C = q_pair[0].shape[2]
    q_all, o_all, do_all = backend.cat_seq(q_pair), backend.cat_seq(o_pair), backend.cat_seq(do_pair)
    l_all = backend.cat_seq(l_pair)
    kv_cur = backend.pack_kv(k_pair, v_pair)             # [2, B, H, 2C, D]
    dq = backend.zeros_f32(q_all)
    dkv_cur = backend.zeros_f32(kv_cur)                  # gradient accumulators of the K / V on hand, fp32
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    for step in range(world):
        k_all, v_all = backend.unpack_kv(kv_cur)
        (q0, qn), (k0, kn), causal = step_plan(rank, world, step)
        sq, sk = slice(q0 * C, (q0 + qn) * C), slice(k0 * C, (k0 + kn) * C)
        dq_p, dk_p, dv_p = backend.backward(q_all[:, :, sq], k_all[:, :, sk], v_all[:, :, sk], o_all[:, :, sq], l_all[:, :, sq],
                                            do_all[:, :, sq], causal, scale)
        dq[:, :, sq] += dq_p
        dkv_cur[0][:, :, sk] += dk_p
        dkv_cur[1][:, :, sk] += dv_p
        if world > 1:
            if step + 1 < world:
                kv_cur = backend.sendrecv(kv_cur, nxt, prv)
            dkv_cur = backend.sendrecv(dkv_cur, nxt, prv)        # after the last step this hop delivers dK / dV to their owner
    backend.finish()
    dk, dv = dkv_cur[0], dkv_cur[1]
    return (dq[:, :, :C], dq[:, :, C:]), (dk[:, :, :C], dk[:, :, C:]), (dv[:, :, :C], dv[:, :, C:])


def block_backward_numpy(q, k, v, o, l, do, causal, scale):
    """fp64 reference of one rectangle of the ring backward: gradients of a (query block, key block) pair given the FINAL L (log2
    units) and O of the query rows.  causal = rows and keys share their local origin (the ring's step 0)."""
    q, k, v, o, l, do = (np.asarray(x, np.float64) for x in (q, k, v, o, l, do))
    s = np.einsum("bhqd,bhkd->bhqk", q, k) * scale
    p = np.exp2(s * LOG2E - l[..., None])
    if causal:
        Sq, Sk = s.shape[-2:]
        p = np.where(np.arange(Sk)[None, :] > np.arange(Sq)[:, None], 0.0, p)
    dterm = (do * o).sum(-1)
    dp = np.einsum("bhqd,bhkd->bhqk", do, v)
    ds = p * (dp - dterm[..., None]) * scale
    return (np.einsum("bhqk,bhkd->bhqd", ds, k).astype(np.float32), np.einsum("bhqk,bhqd->bhkd", ds, q).astype(np.float32),
            np.einsum("bhqk,bhqd->bhkd", p, do).astype(np.float32))


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

    def cat_seq(self, pair):
        return np.ascontiguousarray(np.concatenate([pair[0], pair[1]], axis=2))

    def pack_kv(self, k_pair, v_pair):
        return np.ascontiguousarray(np.stack([self.cat_seq(k_pair), self.cat_seq(v_pair)]))

    def unpack_kv(self, buf):
        return buf[0], buf[1]

    def attend(self, q, k, v, causal, scale):
        o, l = self.attend_fn(q, k, v, causal, scale)
        return [np.array(o, np.float32), np.array(l, np.float32)]

    def attend_accumulate(self, acc, row0, q, k, v, causal, scale):
        o, l = self.attend_fn(q, k, v, causal, scale)
        rows = q.shape[2]
        o_acc = np.ascontiguousarray(acc[0][:, :, row0:row0 + rows])
        l_acc = np.ascontiguousarray(acc[1][:, :, row0:row0 + rows])
        merge_partials_numpy(o_acc, l_acc, np.asarray(o, np.float32), np.asarray(l, np.float32))
        acc[0][:, :, row0:row0 + rows] = o_acc
        acc[1][:, :, row0:row0 + rows] = l_acc

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

    def zeros_f32(self, like):
        return np.zeros(like.shape, np.float32)

    def backward(self, q, k, v, o, l, do, causal, scale):
        return block_backward_numpy(q, k, v, o, l, do, causal, scale)

    def sendrecv(self, buf, dst, src):
        import torch
        send = torch.from_numpy(np.ascontiguousarray(buf))
        recv = torch.empty_like(send)
        for r in [self.dist.isend(send, dst), self.dist.irecv(recv, src)]:
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
        self.comm = torch.cuda.Stream(device, priority=-1)
        self.stream_ptr = ctypes.c_void_p(self.compute.cuda_stream)
        self.kv_bufs = [None, None]
        self.compute_done = [None, None]
        self.scratch = {}
        self.launches = 0

    def _buf(self, t):
        from .core import MFABuffer
        return MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size())

    def _view(self, t):
        """handle over a (possibly sequence-sliced) [B, H, S, D] tensor: element strides travel with the handle"""
        from .core import MFABuffer
        span = sum((n - 1) * st for n, st in zip(t.shape, t.stride())) + 1
        return MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=span * t.element_size(), shape=tuple(t.shape),
                         strides=tuple(t.stride()))

    def cat_seq(self, pair):
        return self.torch.cat([pair[0], pair[1]], dim=2)

    def pack_kv(self, k_pair, v_pair):
        torch = self.torch
        buf = torch.stack([self.cat_seq(k_pair), self.cat_seq(v_pair)]).contiguous()
        self.kv_bufs[0] = buf
        self.kv_bufs[1] = torch.empty_like(buf)
        self.cur = 0
        return buf

    def unpack_kv(self, buf):
        return buf[0], buf[1]

    def attend(self, q, k, v, causal, scale):
        torch = self.torch
        B, H, S, D = q.shape
        Skv = k.shape[2]
        o = torch.empty(B, H, S, D, device=self.device, dtype=torch.float32)
        l = torch.empty(B, H, S, device=self.device, dtype=torch.float32)
        bufs = [self._view(q), self._view(k), self._view(v), self._buf(o), self._buf(l)]
        rc = self.lib.mfa_attention_forward_ex(self.ctx.handle, *[b.handle for b in bufs], B, S, Skv, H, D, scale, causal,
                                               -1, self.prec, 2, None, 0, None, None, 0, 0, 0, self.stream_ptr)
        for b in bufs:
            b.close()
        if rc != 0:
            raise RuntimeError(f"mfa_attention_forward_ex failed: {rc}")
        self.launches += 1
        return [o, l]

    def attend_accumulate(self, acc, row0, q, k, v, causal, scale):
        B, H, S, D = q.shape
        Skv = k.shape[2]
        T = acc[1].shape[2]
        bufs = [self._view(q), self._view(k), self._view(v), self._buf(acc[0]), self._buf(acc[1])]
        rc = self.lib.mfa_attention_forward_accumulate(self.ctx.handle, *[b.handle for b in bufs], B, S, Skv, H, D, scale,
                                                       causal, -1, self.prec, row0, T, self.stream_ptr)
        for b in bufs:
            b.close()
        if rc != 0:
            raise RuntimeError(f"mfa_attention_forward_accumulate failed: {rc}")
        self.launches += 1

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

    def zeros_f32(self, like):
        return self.torch.zeros(like.shape, device=self.device, dtype=self.torch.float32)

    def backward(self, q, k, v, o, l, do, causal, scale):
        """one rectangle through mfa_attention_backward_ex (tensor-core dK/dV + dQ kernels); windows are made contiguous"""
        torch = self.torch
        q, k, v, do = (t.contiguous() for t in (q, k, v, do))
        o, l = o.contiguous(), l.contiguous()
        B, H, S, D = q.shape
        Skv = k.shape[2]
        dq = torch.empty(B, H, S, D, device=self.device, dtype=torch.float32)
        dk = torch.empty(B, H, Skv, D, device=self.device, dtype=torch.float32)
        dv = torch.empty(B, H, Skv, D, device=self.device, dtype=torch.float32)
        bufs = [self._buf(t) for t in (do, q, k, v, o, l, dq, dk, dv)]
        rc = self.lib.mfa_attention_backward_ex(self.ctx.handle, *[b.handle for b in bufs], None, B, S, Skv, H, D, scale, causal,
                                                -1, self.prec, None, 0, None, None, 0, 0, 0, self.stream_ptr)
        for b in bufs:
            b.close()
        if rc != 0:
            raise RuntimeError(f"mfa_attention_backward_ex failed: {rc}")
        self.launches += 1
        return dq, dk, dv

    def sendrecv(self, buf, dst, src):
        torch, dist = self.torch, self.dist
        buf = buf.contiguous()
        recv = torch.empty_like(buf)
        for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, dst), dist.P2POp(dist.irecv, recv, src)]):
            r.wait()
        return recv

    def finish(self):
        pass


# ------------------------------------------------------------------------------------------------ runners (bench / adapters)
class PythonRingRunner:
    """Ring forward driven from Python: `ring_attention_forward` over `CudaBackend` (torch.distributed P2P for the hops)."""
    kind = "python driver (umfa/ring.py) over mfa_attention_forward_accumulate"
    transport = "NCCL send/recv (torch.distributed.batch_isend_irecv)"

    def __init__(self, ctx, dist, device, dtype, rank, world):
        self.be = CudaBackend(ctx, dist if world > 1 else None, device, dtype)
        self.rank, self.world = rank, world

    @property
    def launches(self):
        return self.be.launches

    def forward(self, q_pair, k_pair, v_pair, scale):
        return ring_attention_forward(self.be, q_pair, k_pair, v_pair, self.rank, self.world, scale)

    def backward(self, q_pair, k_pair, v_pair, o_pair, l_pair, do_pair, scale):
        return ring_attention_backward(self.be, q_pair, k_pair, v_pair, o_pair, l_pair, do_pair, self.rank, self.world, scale)

    def pack(self, q_pair, k_pair, v_pair):
        return (q_pair, k_pair, v_pair)

    def forward_packed(self, pk, scale):
        return self.forward(*pk, scale)

    def close(self):
        pass


def make_runner(ctx, dist, device, dtype, rank, world, prefer_native=True):
    """The ring driver bench.py and the adapters use: the native one (libMFAFFI's mfa_ring_* symbols: C++ step loop,
    ncclSend / ncclRecv on a side stream) when the library has it and NCCL can be loaded, else the Python driver."""
    if prefer_native:
        try:
            from .ring_native import NativeRingRunner
            return NativeRingRunner(ctx, dist, device, dtype, rank, world)
        except Exception:
            pass
    return PythonRingRunner(ctx, dist, device, dtype, rank, world)
