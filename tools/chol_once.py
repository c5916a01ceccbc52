"""One Cholesky (device API) of N x N after a warm-up run: for NAB_CHOL_TRACE=1 timelines. Usage: python tools/chol_once.py [N]"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A0 = torch.empty(N * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
_capi.check(L.na_fill_spd_block_dev(A0.data_ptr(), N, N, N, 5, 0, 0, N, s))
fail = C.c_size_t(0)
for it in range(2):
    A.copy_(A0); torch.cuda.synchronize()
    if it == 1: print("==== timed run", file=sys.stderr, flush=True)
    _capi.check(L.na_cholesky_f64_dev(N, A.data_ptr(), N, 0, 0.0, C.addressof(fail), s)); torch.cuda.synchronize()
