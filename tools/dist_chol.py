"""torchrun tool: block-cyclic Cholesky timing.  torchrun --nproc-per-node P tools/dist_chol.py N NB"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from nalgebra_b200.distributed import ColumnBlockCyclic, DeviceOps, cholesky_block_cyclic, lu_block_cyclic
from nalgebra_b200 import _capi
n = int(sys.argv[1]); nb = int(sys.argv[2]); la = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
mode = sys.argv[4] if len(sys.argv) > 4 else "chol"
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
else: dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29641", rank=0, world_size=1)
A = ColumnBlockCyclic(n, nb, rank, world, DeviceOps(torch.device(f"cuda:{lr}")))
def fill_lu():
    s = torch.cuda.current_stream().cuda_stream
    for b in A.my_blocks:
        _capi.check(A.ops.lib.na_fill_uniform_block_dev(A.ptr(0, b), n, A.width(b), n, 6, 0, b * nb, n, s))
for it in range(3):
    (A.fill_spd(5) if mode == "chol" else fill_lu()); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    st = cholesky_block_cyclic(A, lookahead=la) if mode == "chol" else len(lu_block_cyclic(A, lookahead=la))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{lr}")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        fl = n**3/3 if mode == "chol" else 2*n**3/3
        print(f"block-cyclic {mode} n={n} nb={nb} world={world} lookahead={la}: status/nswaps {st} {t.item()*1e3:.1f} ms  {fl/t.item()/1e12:.2f} TFLOP/s aggregate ({fl/t.item()/1e12/37.18/world*100:.1f}% of {world}x peak)", flush=True)
dist.destroy_process_group()
