"""Times nn_gemm128_chain against the two single launches it replaces (B200; M = c2's pair count)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from newtonnet_b200 import _lib as L
from test_gpu_chain import _chain
from test_gpu_parity import _gemm

lib = L.load(); lib.nn_set_gemm_backend(2)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_802_624
dev = 'cuda:0'
X, aux, acc = (torch.randn(M, 128, device=dev) for _ in range(3))
B1, B2 = torch.randn(128, 128, device=dev) / 11, torch.randn(128, 128, device=dev) / 11
mid, Y = torch.empty_like(X), torch.empty_like(X)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def two_fwd():
    _gemm(lib, X, B1, Y=mid); _gemm(lib, mid, B2, pro=L.PRO_SILU_SAVE, aux_out=mid, Y=Y)


def two_bwd(add):
    _gemm(lib, X, B1, epi=L.EPI_MUL, aux1=aux, Y=mid)
    _gemm(lib, mid, B2, epi=L.EPI_ADD if add else L.EPI_BIAS, aux1=acc if add else None, Y=acc if add else Y)


print('M', M)
print('fwd  two launches %.3f ms   chain %.3f ms' % (timeit(two_fwd), timeit(lambda: _chain(lib, X, B1, B2, 0, 0, aux_out=mid, Y=Y))))
print('bwd  two launches %.3f ms   chain %.3f ms' % (timeit(lambda: two_bwd(False)), timeit(lambda: _chain(lib, X, B1, B2, 1, 0, aux1=aux, Y=Y))))
print('bwd+ two launches %.3f ms   chain %.3f ms' % (timeit(lambda: two_bwd(True)), timeit(lambda: _chain(lib, X, B1, B2, 1, 1, aux1=aux, aux2=acc, Y=acc))))
