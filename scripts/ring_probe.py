"""Times the pieces of one ring step (NCCL hop alone, attention alone, both overlapped) under torchrun."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); os.environ["MFA_CUDA_DEVICE"] = str(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import umfa
from umfa import ring
dev = torch.device("cuda", lr)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C = N // (2 * world); D = 128
ctx = umfa.MFAContext()
be = ring.CudaBackend(ctx, dist, dev, "bf16")
mk = lambda: torch.randn(1, H, C, D, device=dev).to(torch.bfloat16)
qp, kp, vp = (mk(), mk()), (mk(), mk()), (mk(), mk())
buf = be.pack_kv(kp, vp)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def hop():
    ops = [dist.P2POp(dist.isend, buf, (rank + 1) % world), dist.P2POp(dist.irecv, be.kv_bufs[1], (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops): r.wait()
def attn():
    be.attend(qp[0], kp[0], vp[0], False, 0.088, out="scratch"); be.attend(qp[1], kp[0], vp[0], False, 0.088, out="scratch")
def full():
    ring.ring_attention_forward(be, qp, kp, vp, rank, world, 0.088)
r = {"hop_ms": t(hop), "attn2_ms": t(attn), "ring_ms": t(full), "hop_MB": buf.numel() * 2 / 1e6}
if rank == 0: print(r, flush=True)
dist.barrier(); dist.destroy_process_group()
