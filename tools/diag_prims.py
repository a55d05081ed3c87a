"""Double-backward check of the training primitives against torch-native fp64 equivalents."""
import sys
import torch
import torch.nn.functional as Fn
sys.path.insert(0, '.')
from newtonnet_b200.train import Gemm, GemmTN, Gather, SegmentSum, Segments, linear
dev = torch.device('cuda:0')
torch.manual_seed(0)

def run(M, use_mine, dt):
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(M, 128, generator=g); W1 = torch.randn(128, 128, generator=g) / 11; b1 = torch.randn(128, generator=g)
    W2 = torch.randn(128, 128, generator=g) / 11; b2 = torch.randn(128, generator=g); tgt = torch.randn(M, 128, generator=g)
    x = x0.to(dev, dt).requires_grad_(True)
    P = [t.to(dev, dt).requires_grad_(True) for t in (W1, b1, W2, b2)]
    lin = (lambda a, w, b: linear(a, w, b)) if use_mine else (lambda a, w, b: Fn.linear(a, w, b))
    y = lin(Fn.silu(lin(x, P[0], P[1])), P[2], P[3])
    e = (y * y).sum()
    gx, = torch.autograd.grad(e, x, create_graph=True)
    loss = ((gx - tgt.to(dev, dt)) ** 2).mean() + e * 1e-3
    return [t.double().cpu() for t in torch.autograd.grad(loss, P)]

for M in (64, 128, 914, 2742):
    ref = run(M, False, torch.float64)
    nat = run(M, False, torch.float32)
    mine = run(M, True, torch.float32)
    print(f'M={M}: native fp32 vs fp64 ' + ' '.join(f'{((a-b).abs().max()/b.abs().max()).item():.1e}' for a, b in zip(nat, ref)) +
          ' | mine vs fp64 ' + ' '.join(f'{((a-b).abs().max()/b.abs().max()).item():.1e}' for a, b in zip(mine, ref)))

# gather / segment-sum double backward
def run2(use_mine, dt):
    g = torch.Generator().manual_seed(2)
    N, E = 300, 4000
    idx = torch.randint(0, N, (E,), generator=g).sort().values.to(dev)
    idx2 = idx[torch.randperm(E, generator=g).to(dev)]
    rows0 = torch.randn(N, 128, generator=g); wv = torch.randn(E, 128, generator=g); tgt = torch.randn(N, 128, generator=g)
    rows = rows0.to(dev, dt).requires_grad_(True); w = wv.to(dev, dt).requires_grad_(True)
    if use_mine:
        s1, s2 = Segments(idx, N), Segments(idx2, N)
        m = Gather.apply(rows, s1) * Gather.apply(rows, s2) * w
        out = SegmentSum.apply(m, s1)
    else:
        m = rows[idx] * rows[idx2] * w
        out = torch.zeros(N, 128, dtype=dt, device=dev).index_add(0, idx, m)
    e = (out ** 3).sum()
    gr, = torch.autograd.grad(e, rows, create_graph=True)
    loss = ((gr - tgt.to(dev, dt)) ** 2).mean()
    return [t.double().cpu() for t in torch.autograd.grad(loss, [rows, w])]
ref = run2(False, torch.float64); nat = run2(False, torch.float32); mine = run2(True, torch.float32)
print('gather/segsum: native fp32 vs fp64 ' + ' '.join(f'{((a-b).abs().max()/b.abs().max()).item():.1e}' for a, b in zip(nat, ref)) +
      ' | mine vs fp64 ' + ' '.join(f'{((a-b).abs().max()/b.abs().max()).item():.1e}' for a, b in zip(mine, ref)))
