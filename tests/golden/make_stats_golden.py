"""Golden statistics from the reference's own MolecularStatistics (newtonnet/data/loader.py:197-230).
Run in the build container only (needs /root/reference): python tests/golden/make_stats_golden.py
ase / torch_geometric are absent here; both are stubbed with the few names loader.py touches at import time -
MolecularStatistics itself only uses torch and torch_geometric.utils.scatter (the stub from make_golden.py)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.argv = [sys.argv[0]]
import make_golden as mg      # installs the torch_geometric.utils.scatter stub, provides the raw xyz reader

tgd = types.ModuleType('torch_geometric.data')
class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)
for name in ('Dataset', 'InMemoryDataset', 'Batch'):
    setattr(tgd, name, type(name, (), {}))
tgd.Data = Data
sys.modules['torch_geometric.data'] = tgd
ase = types.ModuleType('ase'); ase.units = types.SimpleNamespace(kcal=1.0, kJ=1.0, mol=1.0)
aseio = types.ModuleType('ase.io'); aseio.read = None; ase.io = aseio
sys.modules['ase'] = ase; sys.modules['ase.io'] = aseio; sys.modules['ase.units'] = ase.units
spec = importlib.util.spec_from_file_location('ref_loader', '/root/reference/newtonnet/data/loader.py')
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)

zs, ps, es, fs = mg.read_extxyz('/root/reference/scripts/md17_data/aspirin/ccsd_train/raw/aspirin_ccsd-train.xyz', 60)
rng = np.random.default_rng(0)
frames = []                   # vary the composition so the least-squares problem is well posed
for k in range(60):
    keep = np.sort(rng.choice(21, size=rng.integers(12, 22), replace=False))
    frames.append(dict(z=zs[k][keep], e=es[k] * len(keep) / 21 + 0.01 * k, f=fs[k][keep]))
z = torch.tensor(np.concatenate([f['z'] for f in frames]))
batch = torch.tensor(np.concatenate([np.full(len(f['z']), b) for b, f in enumerate(frames)]))
data = Data(z=z, batch=batch, energy=torch.tensor([f['e'] for f in frames], dtype=torch.float64),
            force=torch.tensor(np.concatenate([f['f'] for f in frames]), dtype=torch.float64))
st = ref.MolecularStatistics()(data)
np.savez_compressed(os.path.join(HERE, 'stats_aspirin60.npz'), z=z.numpy(), batch=batch.numpy(), energy=data.energy.numpy(),
                    force=data.force.numpy(), e_shift=st['energy']['shift'].numpy(), e_scale=st['energy']['scale'].numpy(),
                    f_scale=st['force']['scale'].numpy())
print('shift[1,6,8]', st['energy']['shift'][[1, 6, 8]].numpy(), 'scale', st['energy']['scale'][6].item(),
      'fscale', st['force']['scale'][[1, 6, 8]].numpy())
