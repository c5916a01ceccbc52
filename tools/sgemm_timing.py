"""f32 GEMM (tcgen05 3xTF32) device timing. Usage: python tools/sgemm_timing.py [N ...]"""
import sys
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
for N in ([int(x) for x in sys.argv[1:]] or [4096, 8192, 16384]):
    A = torch.rand(N * N, dtype=torch.float32, device=dev) - 0.5; B = torch.rand(N * N, dtype=torch.float32, device=dev) - 0.5
    C = torch.empty(N * N, dtype=torch.float32, device=dev)
    for _ in range(2):
        _capi.check(L.na_sgemm_dev(N, N, N, 1.0, A.data_ptr(), 1, N, B.data_ptr(), 1, N, 0.0, C.data_ptr(), 1, N, s))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        _capi.check(L.na_sgemm_dev(N, N, N, 1.0, A.data_ptr(), 1, N, B.data_ptr(), 1, N, 0.0, C.data_ptr(), 1, N, s))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ref = (A.view(N, N).t()[:64].double() @ B.view(N, N).t().double()); got = C.view(N, N).t()[:64].double()
    print(f"sgemm {N}^3: {ms:8.3f} ms  {2*N**3/ms/1e9:8.1f} TFLOP/s (f32-equivalent; 3x that many TF32 flops)  max err vs f64 {(got-ref).abs().max().item():.2e}", flush=True)
