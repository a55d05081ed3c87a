#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box).
The reference (THGLab/NewtonNet v2.1.0) is imported from where it lies; the two third-party
modules its hot path imports but that are not installed here are replaced by in-memory stubs:

  * torch_geometric.utils.scatter  (call sites newtonnet/models/newtonnet.py:214,226 and
    newtonnet/models/output.py:235,246) = src.new_zeros(size).scatter_add_(dim, index, src)
  * les.Les (newtonnet/models/output.py:5,207-212,229-231) - only constructed, never called,
    on the energy / gradient_force / stress / virial heads.

Outputs (all small, committed):
  md17_kat.npz        201 frames of scripts/md17_md/md.traj (positions, energy, forces) - the
                      reference's own calculator output (fp32, CUDA) - plus numbers.
  weights_md17.npz    shipped checkpoint scripts/md17_model/training_1/models/best_model.pt as a
                      flat fp32 state_dict with the current key names.
  weights_seed0.npz   default-init NewtonNet (torch.manual_seed(0)) with randomised scale/shift.
  case_*.npz          inputs + reference outputs (fp32 run and fp64 run) for each parity case.

Usage: python tests/golden/make_golden.py
"""
import json
import os
import struct
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- stubs
def _install_stubs():
    tg = types.ModuleType('torch_geometric')
    tgu = types.ModuleType('torch_geometric.utils')

    def scatter(src, index, dim=0, dim_size=None, reduce='sum'):
        if dim < 0:
            dim += src.dim()
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        shape = list(src.shape)
        shape[dim] = dim_size
        view = [1] * src.dim()
        view[dim] = -1
        idx = index.view(view).expand_as(src)
        out = src.new_zeros(shape).scatter_add_(dim, idx, src)
        if reduce in ('sum', 'add'):
            return out
        if reduce == 'mean':
            cnt = src.new_zeros(shape).scatter_add_(dim, idx, torch.ones_like(src)).clamp(min=1)
            return out / cnt
        raise NotImplementedError(reduce)

    tgu.scatter = scatter
    tg.utils = tgu
    sys.modules['torch_geometric'] = tg
    sys.modules['torch_geometric.utils'] = tgu

    les = types.ModuleType('les')

    class Les(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.atomwise = torch.nn.Identity()
            self.ewald = torch.nn.Identity()
            self.bec = torch.nn.Identity()

    les.Les = Les
    sys.modules['les'] = les


_install_stubs()
sys.path.insert(0, REF)
from newtonnet.models.newtonnet import NewtonNet  # noqa: E402
import newtonnet.models.output as ref_output  # noqa: E402


# --------------------------------------------------------------------------- fixtures from the repo
def read_ulm_traj(path):
    """ASE ULM v3 reader (no ASE): header, int64 offset table, JSON items + raw ndarrays."""
    buf = open(path, 'rb').read()
    assert buf[:8] == b'- of Ulm'
    _, nitems, pos0 = struct.unpack('<qqq', buf[24:48])
    offs = np.frombuffer(buf[pos0:pos0 + 8 * nitems], dtype='<i8')

    def arr(spec):
        shape, dtype, off = spec['ndarray']
        n = int(np.prod(shape))
        return np.frombuffer(buf, dtype=np.dtype(dtype), count=n, offset=off).reshape(shape).copy()

    frames = []
    numbers = None
    for o in offs:
        n = struct.unpack('<q', buf[o:o + 8])[0]
        item = json.loads(buf[o + 8:o + 8 + n])
        if 'numbers.' in item:
            numbers = arr(item['numbers.'])
        calc = item['calculator.']
        frames.append((arr(item['positions.']), calc['energy'], arr(calc['forces.'])))
    pos = np.stack([f[0] for f in frames])
    en = np.array([f[1] for f in frames])
    frc = np.stack([f[2] for f in frames])
    return numbers, pos, en, frc


def read_extxyz(path, nframes):
    sym = {'H': 1, 'C': 6, 'N': 7, 'O': 8}
    zs, ps, es, fs = [], [], [], []
    with open(path) as fh:
        for _ in range(nframes):
            n = int(fh.readline())
            hdr = fh.readline()
            e = float(hdr.split('energy=')[1].split()[0])
            z, p, f = [], [], []
            for _ in range(n):
                t = fh.readline().split()
                z.append(sym[t[0]])
                p.append([float(x) for x in t[1:4]])
                f.append([float(x) for x in t[4:7]])
            zs.append(z); ps.append(p); es.append(e); fs.append(f)
    return np.array(zs), np.array(ps), np.array(es), np.array(fs)


def load_shipped_state_dict():
    """Appendix-A recipe of SURVEY.md: shim the legacy SumAggregator class, rename two keys."""
    class SumAggregator(torch.nn.Module):
        pass
    ref_output.SumAggregator = SumAggregator
    m = torch.load(f'{REF}/scripts/md17_model/training_1/models/best_model.pt',
                   map_location='cpu', weights_only=False)
    sd = m.state_dict()
    ren = {
        'embedding_layer.node_embedding.weight': 'embedding_layers.node_embedding.weight',
        'embedding_layer.edge_embedding.frequencies': 'embedding_layers.edge_embedding.embedding.frequencies',
    }
    return {ren.get(k, k): v for k, v in sd.items()}


# --------------------------------------------------------------------------- synthetic generators
def water_box(nside, seed=0, dtype=torch.float32):
    """synthetic water lattice (O,H,H order), cubic cell L = nside*3.104."""
    g = torch.Generator().manual_seed(seed)
    a = 3.104
    idx = torch.stack(torch.meshgrid(*[torch.arange(nside)] * 3, indexing='ij'), -1).reshape(-1, 3).to(dtype)
    n = idx.shape[0]
    O = (idx + 0.5) * a + (torch.rand(n, 3, generator=g, dtype=dtype) - 0.5) * 0.4
    Q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=dtype))
    ang = np.deg2rad(104.52)
    h1 = torch.tensor([0.9572, 0.0, 0.0], dtype=dtype)
    h2 = torch.tensor([0.9572 * np.cos(ang), 0.9572 * np.sin(ang), 0.0], dtype=dtype)
    H1 = O + Q @ h1
    H2 = O + Q @ h2
    L = nside * a
    pos = torch.stack([O, H1, H2], 1).reshape(-1, 3) % L
    z = torch.tensor([8, 1, 1]).repeat(n)
    cell = torch.eye(3, dtype=dtype).unsqueeze(0) * L
    batch = torch.zeros(3 * n, dtype=torch.long)
    return z, pos.to(dtype), cell, batch


def molecule_batch(n_mol, seed=1, lo=4, hi=65, dtype=torch.float32, sizes=None):
    """n distinct sites of a (ceil(n^(1/3))+1)^3 lattice, 1.4 A spacing, +-0.25 A jitter."""
    g = torch.Generator().manual_seed(seed)
    if sizes is None:
        sizes = torch.randint(lo, hi, (n_mol,), generator=g).tolist()
    zs, ps, bs = [], [], []
    for b, n in enumerate(sizes):
        m = int(np.ceil(n ** (1 / 3))) + 1
        sites = torch.randperm(m ** 3, generator=g)[:n]
        ijk = torch.stack([sites // (m * m), (sites // m) % m, sites % m], 1).to(dtype)
        p = ijk * 1.4 + (torch.rand(n, 3, generator=g, dtype=dtype) - 0.5) * 0.5
        zs.append(torch.tensor([1, 6, 7, 8])[torch.randint(0, 4, (n,), generator=g)])
        ps.append(p)
        bs.append(torch.full((n,), b, dtype=torch.long))
    z = torch.cat(zs); pos = torch.cat(ps); batch = torch.cat(bs)
    cell = torch.zeros(len(sizes), 3, 3, dtype=dtype)
    return z, pos, cell, batch


# --------------------------------------------------------------------------- running the reference
def build_model(sd, props, dtype):
    m = NewtonNet(output_properties=list(props))
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    m = m.to(dtype)
    m.eval()
    return m


def run_reference(sd, props, z, pos, cell, batch, dtype, want_layers=False):
    m = build_model(sd, props, dtype)
    p = pos.to(dtype).clone().detach()
    c = cell.to(dtype).clone()
    out = m(z, p, c, batch)
    res = {'energy': out.energy.detach().numpy(),
           'forces': out.gradient_force.detach().numpy(),
           'edge_index': out.edge_index.numpy().astype(np.int32)}
    if 'stress' in props:
        res['stress'] = out.stress.detach().numpy()
    if 'virial' in props:
        res['virial'] = out.virial.detach().numpy()
    if want_layers:
        res['atom_node'] = out.atom_node.detach().numpy()
        res['force_node'] = out.force_node.detach().numpy()
    return res


def save_case(name, sd_name, sd, props, z, pos, cell, batch, want_layers=False):
    r32 = run_reference(sd, props, z, pos, cell, batch, torch.float32, want_layers)
    r64 = run_reference(sd, props, z, pos, cell, batch, torch.float64, want_layers)
    d = dict(weights=sd_name, props=np.array(props), z=z.numpy().astype(np.int64),
             pos=pos.numpy().astype(np.float32), cell=cell.numpy().astype(np.float32),
             batch=batch.numpy().astype(np.int64))
    for k, v in r32.items():
        d[f'ref32_{k}'] = v
    for k, v in r64.items():
        if k == 'edge_index':
            continue
        d[f'ref64_{k}'] = v
    np.savez_compressed(f'{OUT}/case_{name}.npz', **d)
    dE = np.abs(r32['energy'] - r64['energy']).max()
    dF = np.abs(r32['forces'] - r64['forces']).max()
    print(f'{name:>22}: N={len(z)} B={cell.shape[0]} E={r32["edge_index"].shape[1]} '
          f'ref32-vs-ref64 dE={dE:.3e} dF={dF:.3e}')


def main():
    torch.manual_seed(0)
    # ---- weights
    sd_md17 = {k: v.to(torch.float32).numpy() for k, v in load_shipped_state_dict().items()}
    np.savez_compressed(f'{OUT}/weights_md17.npz', **sd_md17)
    torch.manual_seed(0)
    m0 = NewtonNet(output_properties=['energy', 'gradient_force'])
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        m0.scalers[0].scale.weight.copy_(torch.rand(119, 1, generator=g) + 0.5)
        m0.scalers[0].shift.weight.copy_(torch.randn(119, 1, generator=g))
    sd_seed0 = {k: v.detach().numpy().copy() for k, v in m0.state_dict().items()}
    np.savez_compressed(f'{OUT}/weights_seed0.npz', **sd_seed0)
    print('state_dict keys:', len(sd_seed0), 'params:',
          sum(int(np.prod(v.shape)) for v in sd_seed0.values()))

    # ---- known-answer test: the reference's own MD trajectory
    numbers, pos, en, frc = read_ulm_traj(f'{REF}/scripts/md17_md/md.traj')
    np.savez_compressed(f'{OUT}/md17_kat.npz', numbers=numbers, positions=pos, energy=en, forces=frc)
    # sanity: current reference code on CPU fp32 reproduces the trajectory
    zz = torch.tensor(numbers).repeat(8)
    pp = torch.tensor(pos[:8].reshape(-1, 3), dtype=torch.float32)
    bb = torch.arange(8).repeat_interleave(21)
    r = run_reference(sd_md17, ['energy', 'gradient_force'], zz, pp, torch.zeros(8, 3, 3), bb, torch.float32)
    print('KAT check: dE', np.abs(r['energy'] - en[:8]).max(), 'dF', np.abs(r['forces'].reshape(8, 21, 3) - frc[:8]).max())

    EF = ['energy', 'gradient_force']
    EFS = ['energy', 'gradient_force', 'stress', 'virial']

    # ---- C1: aspirin 21 x 100, shipped weights (first 100 test frames)
    zs, ps, es, fs = read_extxyz(f'{REF}/scripts/md17_data/aspirin/ccsd_test/raw/aspirin_ccsd-test.xyz', 100)
    z = torch.tensor(zs.reshape(-1)); p = torch.tensor(ps.reshape(-1, 3), dtype=torch.float32)
    b = torch.arange(100).repeat_interleave(21)
    save_case('aspirin100', 'md17', sd_md17, EF, z, p, torch.zeros(100, 3, 3), b)
    np.savez_compressed(f'{OUT}/aspirin_ccsd_labels.npz', energy=es, forces=fs)
    # single molecule with per-layer outputs
    save_case('aspirin1', 'md17', sd_md17, EF, z[:21], p[:21], torch.zeros(1, 3, 3), b[:21], want_layers=True)

    # ---- periodic water boxes, random weights, with stress
    for nside, nm in [(5, 'water375'), (3, 'water81_smallL'), (7, 'water1029')]:
        z, p, c, b = water_box(nside)
        save_case(nm, 'seed0', sd_seed0, EFS, z, p, c, b, want_layers=(nside == 5))
    # orthorhombic (non-cubic) cell, unwrapped positions (atoms displaced by lattice vectors)
    z, p, c, b = water_box(4)
    c = c.clone(); c[0, 1, 1] *= 1.25; c[0, 2, 2] *= 1.6
    g = torch.Generator().manual_seed(5)
    p = p + torch.randint(-2, 3, p.shape, generator=g).float() * torch.diagonal(c[0])
    save_case('water192_ortho_unwrapped', 'seed0', sd_seed0, EFS, z, p, c, b)
    # two periodic systems in one batch (different cubic cells)
    z1, p1, c1, b1 = water_box(4, seed=2)
    z2, p2, c2, b2 = water_box(5, seed=3)
    save_case('water_batch2', 'seed0', sd_seed0, EFS, torch.cat([z1, z2]), torch.cat([p1, p2]),
              torch.cat([c1, c2]), torch.cat([b1, b2 + 1]))
    # triclinic (non-symmetric) cell: reference quirk (shift uses cell @ n), parity unpinned in SURVEY
    z, p, c, b = water_box(4, seed=4)
    c = c.clone(); c[0, 1, 0] = 2.0; c[0, 2, 0] = -1.5; c[0, 2, 1] = 1.0
    save_case('water192_triclinic', 'seed0', sd_seed0, EFS, z, p, c, b)

    # ---- ragged non-periodic molecule batches (C2-shaped), incl. 1- and 2-atom molecules
    z, p, c, b = molecule_batch(24, seed=1)
    save_case('mols24', 'seed0', sd_seed0, EF, z, p, c, b)
    z, p, c, b = molecule_batch(0, seed=7, sizes=[1, 2, 64, 3, 1, 17, 64, 5])
    save_case('mols_edge', 'seed0', sd_seed0, EF, z, p, c, b)
    z, p, c, b = molecule_batch(256, seed=1)
    save_case('mols256', 'seed0', sd_seed0, EF, z, p, c, b)


def training_golden():
    """Row T: one training step's loss and parameter gradients from the unmodified reference in train mode
    (create_graph=True on the force head, reference models/newtonnet.py:106-113; loss = MSE(E) + w MSE(F),
    train/loss.py:48,104-138; weights 1 / 50, scripts/config.yml:46-51), fp64."""
    sd_seed0 = dict(np.load(f'{OUT}/weights_seed0.npz'))
    for name, (z, p, c, b) in {'mols24': molecule_batch(24, seed=1), 'water81': water_box(3)}.items():
        m = build_model(sd_seed0, ['energy', 'gradient_force'], torch.float64)
        m.train()
        g = torch.Generator().manual_seed(11)
        e_t = torch.randn(c.shape[0], generator=g, dtype=torch.float64)
        f_t = torch.randn(p.shape[0], 3, generator=g, dtype=torch.float64)
        out = m(z, p.double().clone(), c.double(), b)
        loss = torch.nn.MSELoss()(out.energy, e_t) + 50.0 * torch.nn.MSELoss()(out.gradient_force, f_t)
        loss.backward()
        d = dict(z=z.numpy(), pos=p.numpy().astype(np.float32), cell=c.numpy().astype(np.float32), batch=b.numpy(),
                 e_target=e_t.numpy(), f_target=f_t.numpy(), loss=loss.item(), force_weight=50.0)
        for k, v in m.named_parameters():
            d['grad.' + k] = np.zeros(v.shape) if v.grad is None else v.grad.numpy()
        np.savez_compressed(f'{OUT}/train_{name}.npz', **d)
        gn = np.sqrt(sum((d[k] ** 2).sum() for k in d if k.startswith('grad.')))
        print(f'train_{name}: loss {loss.item():.6f} |grad| {gn:.4f}')


def widening_golden():
    """layer_norm=True and the direct_force head.  Weights = weights_seed0 plus small
    extra tensors stored in the case file (layer-norm affine parameters / direct-force head)."""
    sd0 = dict(np.load(f'{OUT}/weights_seed0.npz'))
    g = torch.Generator().manual_seed(77)
    extra = {}
    for l in range(3):
        extra[f'interaction_layers.{l}.layer_norm.weight'] = (1.0 + 0.3 * torch.randn(128, generator=g)).numpy()
        extra[f'interaction_layers.{l}.layer_norm.bias'] = (0.2 * torch.randn(128, generator=g)).numpy()
    torch.manual_seed(3)
    probe = NewtonNet(output_properties=['energy', 'gradient_force', 'direct_force'])
    for k, v in probe.state_dict().items():
        if k.startswith('output_layers.2.') or k.startswith('scalers.2.'):
            extra[k] = v.detach().numpy().copy()
    extra['scalers.2.scale.weight'] = (torch.rand(119, 1, generator=g) + 0.5).numpy()
    for name, (z, p, c, b) in {'mols_edge': molecule_batch(0, seed=7, sizes=[1, 2, 64, 3, 1, 17, 64, 5]),
                               'water81': water_box(3)}.items():
        for dtype, tag in ((torch.float64, 'ref64'), (torch.float32, 'ref32')):
            m = NewtonNet(layer_norm=True, output_properties=['energy', 'gradient_force', 'direct_force'])
            sd = {k: torch.as_tensor(v) for k, v in {**sd0, **extra}.items()}
            m.load_state_dict(sd, strict=True)
            m = m.to(dtype); m.eval()
            out = m(z, p.to(dtype).clone(), c.to(dtype), b)
            if tag == 'ref64':
                d = dict(z=z.numpy(), pos=p.numpy().astype(np.float32), cell=c.numpy().astype(np.float32), batch=b.numpy())
                d.update({'extra.' + k: v for k, v in extra.items()})
            d[f'{tag}_energy'] = out.energy.detach().numpy()
            d[f'{tag}_forces'] = out.gradient_force.detach().numpy()
            d[f'{tag}_direct_force'] = out.direct_force.detach().numpy()
            d[f'{tag}_atom_node'] = out.atom_node.detach().numpy().astype(np.float32)
        np.savez_compressed(f'{OUT}/wide_{name}.npz', **d)
        print(f'wide_{name}: E {d["ref64_energy"][:2]} |F|max {np.abs(d["ref64_forces"]).max():.3f} '
              f'|DF|max {np.abs(d["ref64_direct_force"]).max():.3f} 32-vs-64 dF {np.abs(d["ref32_forces"]-d["ref64_forces"]).max():.2e}')


def hessian_golden():
    """HessianOutput (models/output.py:134-152) of one aspirin frame with the shipped weights, fp64."""
    sd = dict(np.load(f'{OUT}/weights_md17.npz'))
    kat = np.load(f'{OUT}/md17_kat.npz')
    z = torch.tensor(kat['numbers']); p = torch.tensor(kat['positions'][0], dtype=torch.float64)
    m = NewtonNet(output_properties=['energy', 'gradient_force', 'hessian'])
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    m = m.double(); m.eval()
    for layer in m.output_layers:                      # what MLAseCalculator.load_model does (ase_interface.py:125-128)
        if hasattr(layer, 'create_graph'):
            layer.create_graph = True
    out = m(z, p.clone(), torch.zeros(1, 3, 3, dtype=torch.float64), torch.zeros(21, dtype=torch.long))
    h = out.hessian.detach().numpy()
    np.savez_compressed(f'{OUT}/hessian_aspirin1.npz', z=z.numpy(), pos=p.numpy().astype(np.float32), hessian=h,
                        forces=out.gradient_force.detach().numpy())
    print('hessian', h.shape, np.abs(h).max(), 'asym', np.abs(h.reshape(63, 63) - h.reshape(63, 63).T).max())


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'hessian':
        hessian_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == 'wide':
        widening_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == 'train':
        training_golden()
    else:
        main()
        training_golden()
        widening_golden()
        hessian_golden()
