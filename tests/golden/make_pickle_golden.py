"""Writes tests/golden/ref_module_pickle.pt: a WHOLE pickled module saved by the unmodified reference with its
current module tree (train/trainer.py:219 `torch.save(model, ...)`), as utils/ase_interface.py:87 loads it, plus
ref_module_pickle.npz with the reference's outputs on one small batch.

The current reference puts a `les.Les` instance inside aggregators.N (models/output.py:229); `les` is not
installed here, so a stand-in package with the same top-level class path (and a nested submodule class, as the real
package has) is written to a temporary directory - the pickle then names `les.Les` / `les.module.Ewald` exactly as a
real checkpoint does.  Run inside the build container:  python tests/golden/make_pickle_golden.py
"""
import os
import sys
import tempfile
import textwrap

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))

tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, 'les', 'module'))
open(os.path.join(tmp, 'les', '__init__.py'), 'w').write(textwrap.dedent('''
    import torch
    from les.module.ewald import Ewald
    class Les(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.atomwise = torch.nn.Identity()
            self.ewald = Ewald()
            self.bec = torch.nn.Identity()
'''))
open(os.path.join(tmp, 'les', 'module', '__init__.py'), 'w').write('')
open(os.path.join(tmp, 'les', 'module', 'ewald.py'), 'w').write(textwrap.dedent('''
    import torch
    class Ewald(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.sigma = 1.0
'''))
os.makedirs(os.path.join(tmp, 'torch_geometric', 'utils'))
open(os.path.join(tmp, 'torch_geometric', '__init__.py'), 'w').write('')
open(os.path.join(tmp, 'torch_geometric', 'utils', '__init__.py'), 'w').write(textwrap.dedent('''
    import torch
    def scatter(src, index, dim=0, dim_size=None, reduce='sum'):
        if dim < 0:
            dim += src.dim()
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        shape = list(src.shape); shape[dim] = dim_size
        view = [1] * src.dim(); view[dim] = -1
        idx = index.view(view).expand_as(src)
        out = src.new_zeros(shape).scatter_add_(dim, idx, src)
        if reduce in ('sum', 'add'):
            return out
        raise NotImplementedError(reduce)
'''))
sys.path[:0] = [tmp, REF]
from newtonnet.models.newtonnet import NewtonNet  # noqa: E402

torch.manual_seed(11)
model = NewtonNet(n_interactions=1, output_properties=['energy', 'gradient_force'])
with torch.no_grad():
    model.scalers[0].scale.weight.uniform_(0.5, 1.5)
    model.scalers[0].shift.weight.normal_()
assert type(model.aggregators[0].les).__module__ == 'les'
torch.save(model, os.path.join(OUT, 'ref_module_pickle.pt'))

rng = np.random.default_rng(3)
n = 9
z = rng.choice([1, 6, 8], n).astype(np.int64)
pos = (rng.random((n, 3)) * 3.0).astype(np.float32)
model.eval()
out = model(torch.tensor(z), torch.tensor(pos), torch.zeros(1, 3, 3), torch.zeros(n, dtype=torch.long))
np.savez(os.path.join(OUT, 'ref_module_pickle.npz'), z=z, pos=pos, energy=out.energy.detach().numpy(),
         forces=out.gradient_force.detach().numpy())
print('written', os.path.getsize(os.path.join(OUT, 'ref_module_pickle.pt')), 'bytes; E =', out.energy.item())
