"""Data side of the training path: extended-xyz reader, batch collation and
the per-element energy / force statistics that initialise the scalers - without ASE or PyG.

Mirrors reference newtonnet/data/loader.py:167-194 (parse_xyz: wrapped positions, zero cell rows for
non-periodic directions, energy, forces) and :197-230 (MolecularStatistics: per-element energy shifts by
least squares on the composition matrix, one residual RMS per atom as scale, mean force norm per element),
and layers/scalers.py:26-31 (set_scaler_by_string).
"""
import re

import numpy as np
import torch

SYMBOLS = ['X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al', 'Si', 'P', 'S', 'Cl', 'Ar', 'K',
           'Ca', 'Sc', 'Ti', 'V', 'Cr', 'Mn', 'Fe', 'Co', 'Ni', 'Cu', 'Zn', 'Ga', 'Ge', 'As', 'Se', 'Br', 'Kr', 'Rb', 'Sr',
           'Y', 'Zr', 'Nb', 'Mo', 'Tc', 'Ru', 'Rh', 'Pd', 'Ag', 'Cd', 'In', 'Sn', 'Sb', 'Te', 'I', 'Xe', 'Cs', 'Ba', 'La',
           'Ce', 'Pr', 'Nd', 'Pm', 'Sm', 'Eu', 'Gd', 'Tb', 'Dy', 'Ho', 'Er', 'Tm', 'Yb', 'Lu', 'Hf', 'Ta', 'W', 'Re', 'Os',
           'Ir', 'Pt', 'Au', 'Hg', 'Tl', 'Pb', 'Bi', 'Po', 'At', 'Rn', 'Fr', 'Ra', 'Ac', 'Th', 'Pa', 'U', 'Np', 'Pu', 'Am',
           'Cm', 'Bk', 'Cf', 'Es', 'Fm', 'Md', 'No', 'Lr', 'Rf', 'Db', 'Sg', 'Bh', 'Hs', 'Mt', 'Ds', 'Rg', 'Cn', 'Nh', 'Fl',
           'Mc', 'Lv', 'Ts', 'Og']
_Z = {s: i for i, s in enumerate(SYMBOLS)}
_KV = re.compile(r'(\w+)=("([^"]*)"|(\S+))')


def read_extxyz(path, limit=None, length_unit=1.0, energy_unit=1.0):
    """Frames of an extended-xyz file as dicts {z [n] int64, pos [n,3], cell [3,3], energy, force [n,3]} (float64).
    Comment line keys used: Lattice, pbc, energy, Properties (species:S:1:pos:R:3[:forces:R:3])."""
    frames = []
    with open(path) as fh:
        while limit is None or len(frames) < limit:
            line = fh.readline()
            if not line.strip():
                break
            n = int(line)
            info = {m.group(1): (m.group(3) if m.group(3) is not None else m.group(4)) for m in _KV.finditer(fh.readline())}
            cols, off, layout = info.get('Properties', 'species:S:1:pos:R:3').split(':'), 0, {}
            for name, _, width in zip(cols[0::3], cols[1::3], cols[2::3]):
                layout[name] = (off, off + int(width))
                off += int(width)
            rows = [fh.readline().split() for _ in range(n)]
            z = np.array([_Z[r[layout['species'][0]]] for r in rows], dtype=np.int64)
            take = lambda name: np.array([[float(x) for x in r[layout[name][0]:layout[name][1]]] for r in rows])
            pos = take('pos')
            force = take('forces') if 'forces' in layout else (take('force') if 'force' in layout else None)
            cell = np.array([float(x) for x in info['Lattice'].split()]).reshape(3, 3) if 'Lattice' in info else np.zeros((3, 3))
            pbc = np.array([t.upper().startswith('T') for t in info.get('pbc', 'F F F').split()]) if 'pbc' in info \
                else np.array(['Lattice' in info] * 3)
            if pbc.any():          # ase get_positions(wrap=True): wrap the periodic fractional coordinates into [0, 1)
                frac = np.linalg.solve(cell.T, pos.T).T
                frac[:, pbc] %= 1.0
                pos = frac @ cell
            cell = cell.copy()
            cell[~pbc] = 0.0
            frames.append({'z': z, 'pos': pos * length_unit, 'cell': cell * length_unit,
                           'energy': float(info['energy']) * energy_unit if 'energy' in info else None,
                           'force': None if force is None else force * energy_unit / length_unit})
    return frames


def write_extxyz(path, frames, append=False):
    """Frames ({z, pos[, cell, energy, force]}) -> extended xyz in the layout read_extxyz / the reference's data files use."""
    with open(path, 'a' if append else 'w') as fh:
        for f in frames:
            z, pos = np.asarray(f['z']), np.asarray(f['pos'], dtype=np.float64)
            force = f.get('force')
            cell = f.get('cell')
            periodic = cell is not None and np.any(np.asarray(cell) != 0)
            head = 'Properties=species:S:1:pos:R:3' + (':forces:R:3' if force is not None else '')
            if f.get('energy') is not None:
                head += ' energy=%.10f' % float(f['energy'])
            if periodic:
                head = 'Lattice="%s" ' % ' '.join('%.10f' % x for x in np.asarray(cell, dtype=np.float64).reshape(-1)) + head + ' pbc="T T T"'
            else:
                head += ' pbc="F F F"'
            fh.write('%d\n%s\n' % (len(z), head))
            for k in range(len(z)):
                row = '%s %.10f %.10f %.10f' % (SYMBOLS[int(z[k])], *pos[k])
                if force is not None:
                    row += ' %.10f %.10f %.10f' % tuple(np.asarray(force, dtype=np.float64)[k])
                fh.write(row + '\n')


def collate(frames, device=None, dtype=torch.float32):
    """Concatenate frames into the (z, pos, cell, batch, energy, force) tensors NewtonNet.forward / training_step take."""
    z = torch.from_numpy(np.concatenate([f['z'] for f in frames]))
    pos = torch.from_numpy(np.concatenate([f['pos'] for f in frames])).to(dtype)
    cell = torch.from_numpy(np.stack([f['cell'] for f in frames])).to(dtype)
    batch = torch.from_numpy(np.concatenate([np.full(len(f['z']), b, dtype=np.int64) for b, f in enumerate(frames)]))
    energy = torch.tensor([f['energy'] for f in frames], dtype=dtype) if frames[0]['energy'] is not None else None
    force = torch.from_numpy(np.concatenate([f['force'] for f in frames])).to(dtype) if frames[0]['force'] is not None else None
    out = [z, pos, cell, batch, energy, force]
    return tuple(t.to(device) if (t is not None and device is not None) else t for t in out)


def molecular_statistics(frames):
    """{'energy': {'shift' [119], 'scale' [119]}, 'force': {'scale' [119]}} as reference data/loader.py:197-230."""
    z = np.concatenate([f['z'] for f in frames])
    batch = np.concatenate([np.full(len(f['z']), b) for b, f in enumerate(frames)])
    z_unique = np.unique(z)
    stats = {}
    if frames[0]['energy'] is not None:
        energy = np.array([f['energy'] for f in frames], dtype=np.float64)
        formula = np.zeros((len(frames), int(z.max()) + 1))
        np.add.at(formula, (batch, z), 1.0)
        solution = np.linalg.lstsq(formula, energy, rcond=None)[0]
        shift = np.zeros(119); shift[z_unique] = solution[z_unique]
        std = np.sqrt(np.square(energy - formula @ solution).sum() / formula.sum())
        scale = np.ones(119); scale[z_unique] = std
        stats['energy'] = {'shift': torch.from_numpy(shift), 'scale': torch.from_numpy(scale)}
    if frames[0]['force'] is not None:
        fnorm = np.linalg.norm(np.concatenate([f['force'] for f in frames]), axis=-1)
        sums = np.bincount(z, weights=fnorm, minlength=119); counts = np.maximum(np.bincount(z, minlength=119), 1)
        fscale = np.ones(119); fscale[z_unique] = (sums / counts)[z_unique]
        stats['force'] = {'scale': torch.from_numpy(fscale)}
    return stats


def fit_scalers(model, stats, fit_scale=True, fit_shift=True):
    """Initialise the model's ScaleShift layers from statistics (reference scripts/newtonnet_train.py:88-90)."""
    from newtonnet_b200.layers.scalers import set_scaler_by_string
    for key, scaler in zip(model.output_properties, model.scalers):
        ref = next(iter(model.parameters()))
        cast = {k: {kk: vv.to(device=ref.device, dtype=ref.dtype) for kk, vv in v.items()} for k, v in stats.items()}
        set_scaler_by_string(key, scaler, cast, fit_scale=fit_scale, fit_shift=fit_shift)
    return model
