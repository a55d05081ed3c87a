"""Diagnostics of the training path: loss / force / gradient errors vs the fp64 reference golden, per backend."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')   # run from the repo root: python tools/diag_train.py [mols24|water81]
from newtonnet_b200 import _lib as L
from newtonnet_b200.compat import model_from_state_dict
from oracle import newtonnet_oracle as O
lib = L.load(); dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'mols24'
d = dict(np.load(f'tests/golden/train_{name}.npz')); w = dict(np.load('tests/golden/weights_seed0.npz'))
ref = O.forward(w, d['z'], d['pos'], d['cell'], d['batch'], dtype=torch.float64)
t = lambda a, dt=None: torch.tensor(a, device=dev, dtype=dt)
for be in (0, 1):
    lib.nn_set_gemm_backend(be)
    m = model_from_state_dict({k: torch.tensor(v) for k, v in w.items()}).to(dev); m.train()
    pos = t(d['pos']).requires_grad_(True)
    out = m(t(d['z']), pos, t(d['cell']), t(d['batch']))
    e = out.energy.detach().cpu().double().numpy(); f = out.gradient_force.detach().cpu().double().numpy()
    loss = torch.nn.functional.mse_loss(out.energy, t(d['e_target'], torch.float32)) + 50.0 * torch.nn.functional.mse_loss(out.gradient_force, t(d['f_target'], torch.float32))
    loss.backward()
    print(f'backend {be}: dE {np.abs(e-ref["energy"]).max():.3e} dF {np.abs(f-ref["forces"]).max():.3e} (|F| max {np.abs(ref["forces"]).max():.2f}) loss rel {abs(loss.item()-float(d["loss"]))/float(d["loss"]):.3e}')
    gmax = max(np.abs(d[k]).max() for k in d if k.startswith('grad.'))
    rows = []
    for k, p in m.named_parameters():
        r = d['grad.' + k]; g = np.zeros(r.shape) if p.grad is None else p.grad.cpu().double().numpy()
        rows.append((np.abs(g - r).max() / max(np.abs(r).max(), 1e-3 * gmax), k, np.abs(r).max()))
    for e_, k, mx in sorted(rows, reverse=True)[:6]:
        print(f'   {e_:.3e}  {k}  (max|ref| {mx:.3e})')
