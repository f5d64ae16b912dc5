"""Native ring attention: thin caller of libMFAFFI.so's mfa_ring_* symbols (csrc/ring.cu).

The C side owns everything that matters -- the NCCL communicator (its own, made from a unique id), the side stream, the visiting
K/V slots, the arrival flags and the single persistent attention launch per forward.  Python only moves the 128-byte unique id
between the ranks (any host channel works; here torch.distributed, which the caller already has) and hands over device pointers.
"""
import ctypes

from . import _ffi
from .core import MFABuffer


class NativeRingRunner:
    kind = "native driver (csrc/ring.cu: mfa_ring_attention_forward, one persistent launch per forward)"
    transport = "NCCL ncclSend/ncclRecv (direct exchange with rank +- s per step)"

    def __init__(self, ctx, dist, device, dtype, rank, world, reserve_sms=None):
        import torch
        self.torch, self.ctx, self.device, self.rank, self.world = torch, ctx, device, rank, world
        self.lib = _ffi._lib
        self.prec = {"bf16": 1, "fp16": 0}[dtype]
        self.handle = ctypes.c_void_p()
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
        if reserve_sms is not None:
            self.lib.mfa_ring_set_reserved_sms(self.handle, int(reserve_sms))
        self.stream = torch.cuda.current_stream(device)

    @property
    def launches(self):
        return int(self.lib.mfa_ring_launch_count(self.handle))

    def pack(self, q_pair, k_pair, v_pair):
        """[low | high] chunk pairs -> the contiguous [B, H, 2C, D] operands the C API takes (do this once, outside timed loops)."""
        torch = self.torch
        q, k, v = (torch.cat([p[0], p[1]], dim=2).contiguous() for p in (q_pair, k_pair, v_pair))
        B, H, T, D = q.shape
        o = torch.empty(B, H, T, D, device=self.device, dtype=torch.float32)
        l = torch.empty(B, H, T, device=self.device, dtype=torch.float32)
        bufs = [MFABuffer(self.ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
        return {"t": (q, k, v, o, l), "b": bufs, "dims": (B, H, T // 2, D)}

    def forward_packed(self, pk, scale):
        B, H, C, D = pk["dims"]
        stream_ptr = ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.mfa_ring_attention_forward(self.handle, *[b.handle for b in pk["b"]], B, C, H, D, scale, self.prec, stream_ptr)
        if rc != 0:
            raise RuntimeError(f"mfa_ring_attention_forward failed: {rc}")
        o, l = pk["t"][3], pk["t"][4]
        return (o[:, :, :C], l[:, :, :C]), (o[:, :, C:], l[:, :, C:])

    def forward(self, q_pair, k_pair, v_pair, scale):
        """Convenience form on chunk pairs: packs (three device copies) on every call; hot loops pack once and call
        forward_packed."""
        pk = self.pack(q_pair, k_pair, v_pair)
        try:
            return self.forward_packed(pk, scale)
        finally:
            for b in pk["b"]:
                b.close()

    def close(self):
        if self.handle:
            self.lib.mfa_ring_destroy(self.handle)
            self.handle = ctypes.c_void_p()
