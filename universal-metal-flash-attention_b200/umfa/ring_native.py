"""Native ring attention: thin caller of libMFAFFI.so's mfa_ring_* symbols (csrc/ring.cu).

The C side owns everything that matters -- the NCCL communicator (its own, made from a unique id), the side stream, the visiting
K/V slots, the arrival / consumed flags, the per-step attention launches.  Python only carries small blobs between the ranks (the
128-byte NCCL id; for the p2p transport the CUDA IPC handles of the slots) over a channel the caller already has
(torch.distributed here; any other works) and hands over device pointers.
"""
import ctypes
import os

from . import _ffi
from .core import MFABuffer


class NativeRingRunner:
    kind = "native driver (csrc/ring.cu: mfa_ring_attention_forward / _backward, one launch per ring step, merge in the epilogue)"

    def __init__(self, ctx, dist, device, dtype, rank, world, transport=None):
        import torch
        self.torch, self.ctx, self.dist, self.device, self.rank, self.world = torch, ctx, dist, device, rank, world
        self.lib = _ffi._lib
        self.prec = {"bf16": 1, "fp16": 0}[dtype]
        self.handle = ctypes.c_void_p()
        self.want = (transport or os.environ.get("MFA_RING_TRANSPORT", "p2p")).lower()      # copy engines unless asked for NCCL
        uid = (ctypes.c_uint8 * 128)()
        if world > 1:
            if not self.lib.mfa_ring_transport_available():
                raise RuntimeError("libnccl could not be loaded by libMFAFFI.so")
            t = torch.zeros(128, dtype=torch.uint8, device=device)
            if rank == 0:
                rc = self.lib.mfa_ring_get_unique_id(ctypes.cast(uid, ctypes.c_void_p), 128)
                if rc != 0:
                    raise RuntimeError(f"mfa_ring_get_unique_id failed: {rc}")
                t.copy_(torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8))
            dist.broadcast(t, src=0)
            ctypes.memmove(uid, bytes(t.cpu().numpy().tobytes()), 128)
        with torch.cuda.device(device):
            rc = self.lib.mfa_ring_create(ctx.handle, ctypes.cast(uid, ctypes.c_void_p), 128, rank, world, ctypes.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"mfa_ring_create failed: {rc}")
        self._p2p_dims = None

    @property
    def transport(self):
        t = int(self.lib.mfa_ring_transport(self.handle))
        return ("NCCL ncclSend/ncclRecv, direct exchange with ranks r +- s" if t == 0 else
                "p2p copy engines (CUDA IPC peer memory + stream memory operations), direct push to rank r + s")

    @property
    def launches(self):
        return int(self.lib.mfa_ring_launch_count(self.handle))

    def _setup_p2p(self, B, C, H, D):
        """prepare -> export -> all-gather of the handle blobs -> import (collective)"""
        torch, lib = self.torch, self.lib
        rc = lib.mfa_ring_prepare(self.handle, B, C, H, D)
        if rc != 0:
            raise RuntimeError(f"mfa_ring_prepare failed: {rc}")
        nb = int(lib.mfa_ring_handle_bytes())
        blob = (ctypes.c_uint8 * nb)()
        rc = lib.mfa_ring_export_handles(self.handle, ctypes.cast(blob, ctypes.c_void_p), nb)
        if rc != 0:
            raise RuntimeError(f"mfa_ring_export_handles failed: {rc}")
        mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).to(self.device)
        allb = [torch.zeros(nb, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        self.dist.all_gather(allb, mine)
        flat = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb)
        buf = (ctypes.c_uint8 * len(flat)).from_buffer_copy(flat)
        rc = lib.mfa_ring_import_handles(self.handle, ctypes.cast(buf, ctypes.c_void_p), nb)
        if rc != 0:
            raise RuntimeError(f"mfa_ring_import_handles failed: {rc}")
        self.dist.barrier()
        self._p2p_dims = (B, C, H, D)

    def pack(self, q_pair, k_pair, v_pair):
        """[low | high] chunk pairs -> the contiguous [B, H, 2C, D] operands the C API takes (do this once, outside timed loops)."""
        torch = self.torch
        q, k, v = (torch.cat([p[0], p[1]], dim=2).contiguous() for p in (q_pair, k_pair, v_pair))
        B, H, T, D = q.shape
        o = torch.empty(B, H, T, D, device=self.device, dtype=torch.float32)
        l = torch.empty(B, H, T, device=self.device, dtype=torch.float32)
        bufs = [MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
        if self.world > 1 and self.want == "p2p" and self._p2p_dims != (B, T // 2, H, D):
            self._setup_p2p(B, T // 2, H, D)
        return {"t": (q, k, v, o, l), "b": bufs, "dims": (B, H, T // 2, D)}

    def forward_packed(self, pk, scale):
        B, H, C, D = pk["dims"]
        stream_ptr = ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.mfa_ring_attention_forward(self.handle, *[b.handle for b in pk["b"]], B, C, H, D, scale, self.prec, stream_ptr)
        if rc != 0:
            raise RuntimeError(f"mfa_ring_attention_forward failed: {rc}")
        o, l = pk["t"][3], pk["t"][4]
        return (o[:, :, :C], l[:, :, :C]), (o[:, :, C:], l[:, :, C:])

    def backward_packed(self, pk, do_pair, scale):
        """mfa_ring_attention_backward on the operands / results of a forward_packed call: do_pair = (low, high) chunks of the
        upstream gradient in the operands' precision.  Returns ((dq_lo, dq_hi), (dk_lo, dk_hi), (dv_lo, dv_hi)), fp32."""
        torch = self.torch
        B, H, C, D = pk["dims"]
        do = torch.cat([do_pair[0], do_pair[1]], dim=2).contiguous()
        grads = [torch.empty(B, H, 2 * C, D, device=self.device, dtype=torch.float32) for _ in range(3)]
        extra = [MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in [do] + grads]
        qb, kb, vb, ob, lb = pk["b"]
        stream_ptr = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        try:
            rc = self.lib.mfa_ring_attention_backward(self.handle, qb.handle, kb.handle, vb.handle, ob.handle, lb.handle,
                                                      extra[0].handle, extra[1].handle, extra[2].handle, extra[3].handle,
                                                      B, C, H, D, scale, self.prec, stream_ptr)
            if rc != 0:
                raise RuntimeError(f"mfa_ring_attention_backward failed: {rc}")
            torch.cuda.current_stream(self.device).synchronize()       # `do` and the handles die with this call
        finally:
            for b in extra:
                b.close()
        return tuple((g[:, :, :C], g[:, :, C:]) for g in grads)

    def forward(self, q_pair, k_pair, v_pair, scale):
        """Convenience form on chunk pairs: packs (three device copies) on every call; hot loops pack once and call
        forward_packed."""
        pk = self.pack(q_pair, k_pair, v_pair)
        try:
            res = self.forward_packed(pk, scale)
            self.torch.cuda.current_stream(self.device).synchronize()      # the packed operands die with this call
            return res
        finally:
            for b in pk["b"]:
                b.close()

    def close(self):
        if self.handle:
            self.torch.cuda.synchronize(self.device)
            self.lib.mfa_ring_destroy(self.handle)
            self.handle = ctypes.c_void_p()
