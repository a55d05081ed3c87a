"""Diagnostics for the tcgen05 GEMM backend (run on a B200): structured inputs whose outputs reveal
operand-layout mistakes, plus accuracy of the 3xTF32 split against fp64."""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
from newtonnet_b200 import _lib as L

lib = L.load()
dev = torch.device('cuda:0')
s = torch.cuda.current_stream().cuda_stream


def gemm(X, B, backend):
    lib.nn_set_gemm_backend(backend)
    a = L.GemmArgs()
    Y = torch.full_like(X, -777.0)
    img = torch.empty(L.NN_B_IMAGE_FLOATS, device=dev)
    L.check(lib.nn_gemm128_prepare_b(B.data_ptr(), img.data_ptr(), s), 'prep')
    a.X, a.B, a.B_img, a.Y, a.m = X.data_ptr(), B.data_ptr(), img.data_ptr(), Y.data_ptr(), X.shape[0]
    L.check(lib.nn_gemm128(C.byref(a), s), 'gemm')
    torch.cuda.synchronize()
    lib.nn_set_gemm_backend(0)
    return Y


k = torch.arange(128, device=dev, dtype=torch.float32)
B = k[:, None] * 128 + k[None, :]          # B[k][n] = 128 k + n  (exact in hi + lo)
X = torch.eye(128, device=dev)
for be_ in (1, 2):
    Y = gemm(X, B, be_)
    print(f'backend {be_} identity @ B: max err', float((Y - B).abs().max()))
if float((Y - B).abs().max()) > 0.5:
    print('Y[0,:8]  ', Y[0, :8].tolist())
    print('Y[1,:8]  ', Y[1, :8].tolist())
    print('Y[8,:8]  ', Y[8, :8].tolist())
    print('Y[:8,0]  ', Y[:8, 0].tolist())
    print('Y[32,:4] ', Y[32, :4].tolist(), ' Y[33,:4]', Y[33, :4].tolist())
    # decode where each output element came from
    src_k = (Y / 128).floor(); src_n = Y - 128 * src_k
    print('row->k map (first 16 rows, col 0):', src_k[:16, 0].tolist())
    print('col->n map (row 0, first 16 cols):', src_n[0, :16].tolist())
for M in (128, 1000, 148 * 128 * 2 + 5):
    g = torch.Generator().manual_seed(M)
    X = torch.randn(M, 128, generator=g).to(dev); B = (torch.randn(128, 128, generator=g) / 11.3).to(dev)
    ref = X.double() @ B.double()
    for be, name in ((0, 'simt'), (1, 'tc  '), (2, 'ts  ')):
        Y = gemm(X, B, be)
        err = (Y.double() - ref).abs().max().item()
        print(f'M={M:6d} {name}: max abs err {err:.3e} (|ref| max {ref.abs().max().item():.2f})')
# timing
M = 1_800_000
X = torch.randn(M, 128, device=dev); B = torch.randn(128, 128, device=dev) / 11.3
AUX = torch.randn(M, 128, device=dev)
for be, name, pro, epi in ((1, 'tc   plain', 0, 0), (2, 'ts   plain', 0, 0), (1, 'tc   silu-save', 3, 0), (2, 'ts   silu-save', 3, 0),
                           (1, 'tc   mul-out', 0, 4), (2, 'ts   mul-out', 0, 4), (1, 'tc   add-out', 0, 2), (2, 'ts   add-out', 0, 2)):
    gemm(X[:1024], B, be)
    lib.nn_set_gemm_backend(be)
    a = L.GemmArgs(); Y = torch.empty_like(X); img = torch.empty(L.NN_B_IMAGE_FLOATS, device=dev)
    lib.nn_gemm128_prepare_b(B.data_ptr(), img.data_ptr(), s)
    a.X, a.B, a.B_img, a.Y, a.m = X.data_ptr(), B.data_ptr(), img.data_ptr(), Y.data_ptr(), M
    a.prologue, a.epilogue, a.aux1 = pro, epi, AUX.data_ptr()
    SAVE = torch.empty_like(X); a.aux_out = SAVE.data_ptr()
    for _ in range(3):
        lib.nn_gemm128(C.byref(a), s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.nn_gemm128(C.byref(a), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name} M={M}: {ms:.3f} ms  {2*M*128*128/ms*1e-9:.1f} TFLOP/s  {2*M*512/ms*1e-6:.0f} GB/s')
lib.nn_set_gemm_backend(0)
