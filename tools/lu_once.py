"""One LU (device API) of N x N, after one warm-up run: for NAB_LU_TRACE=1 timelines. Usage: python tools/lu_once.py [N]"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from nalgebra_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
_capi.check(L.na_init(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A0 = torch.empty(N * N, dtype=torch.float64, device=dev); A = torch.empty_like(A0)
_capi.check(L.na_fill_uniform_dev(A0.data_ptr(), N, N, N, 6, s))
swaps = (C.c_size_t * (2 * N))(); ns = C.c_size_t(0)
for it in range(2):
    A.copy_(A0); torch.cuda.synchronize()
    if it == 1: print("==== timed run", file=sys.stderr, flush=True)
    _capi.check(L.na_lu_f64_dev(N, N, A.data_ptr(), N, swaps, C.addressof(ns), s)); torch.cuda.synchronize()
