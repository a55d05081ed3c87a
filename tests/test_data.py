"""Data side of the training path (newtonnet_b200/data.py) against the reference's statistics golden and its own
writer/reader round trip.  CPU only."""
import os

import numpy as np
import pytest
import torch

from newtonnet_b200 import data as D

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _frames_from_golden():
    g = np.load(os.path.join(GOLDEN, 'stats_aspirin60.npz'))
    frames = []
    for b in range(int(g['batch'].max()) + 1):
        sel = g['batch'] == b
        frames.append({'z': g['z'][sel], 'pos': np.zeros((sel.sum(), 3)), 'cell': np.zeros((3, 3)),
                       'energy': float(g['energy'][b]), 'force': g['force'][sel]})
    return g, frames


def test_statistics_match_reference_golden():
    # reference newtonnet/data/loader.py:197-230 run on these exact arrays (tests/golden/make_stats_golden.py)
    g, frames = _frames_from_golden()
    st = D.molecular_statistics(frames)
    np.testing.assert_allclose(st['energy']['shift'].numpy(), g['e_shift'], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(st['energy']['scale'].numpy(), g['e_scale'], rtol=1e-6)
    np.testing.assert_allclose(st['force']['scale'].numpy(), g['f_scale'], rtol=1e-6)
    assert st['energy']['shift'].shape == (119,)


def _write_extxyz(path, frames, periodic):
    with open(path, 'w') as fh:
        for f in frames:
            fh.write(f"{len(f['z'])}\n")
            head = 'Properties=species:S:1:pos:R:3:forces:R:3 energy=%.10f ' % f['energy']
            if periodic:
                head += 'Lattice="%s" pbc="T T T"' % ' '.join('%.8f' % x for x in f['cell'].reshape(-1))
            else:
                head += 'pbc="F F F"'
            fh.write(head + '\n')
            for zi, p, fo in zip(f['z'], f['pos'], f['force']):
                fh.write('%s %.10f %.10f %.10f %.10f %.10f %.10f\n' % (D.SYMBOLS[zi], *p, *fo))


@pytest.mark.parametrize('periodic', [False, True])
def test_extxyz_round_trip(tmp_path, periodic):
    rng = np.random.default_rng(1)
    cell = np.array([[6.0, 0, 0], [0.5, 7.0, 0], [0, 0.3, 8.0]])
    frames = []
    for k in range(4):
        n = 3 + k
        frames.append({'z': rng.choice([1, 6, 8, 17], size=n), 'pos': rng.uniform(-3, 12, (n, 3)), 'cell': cell,
                       'energy': float(rng.normal()), 'force': rng.normal(size=(n, 3))})
    path = tmp_path / 'frames.xyz'
    _write_extxyz(path, frames, periodic)
    got = D.read_extxyz(str(path))
    assert len(got) == 4 and len(D.read_extxyz(str(path), limit=2)) == 2
    for f, g in zip(frames, got):
        assert (f['z'] == g['z']).all()
        np.testing.assert_allclose(g['force'], f['force'], atol=1e-9)
        assert abs(g['energy'] - f['energy']) < 1e-9
        if periodic:                       # wrapped into the cell: same point modulo lattice vectors, fractional in [0, 1)
            frac = np.linalg.solve(cell.T, g['pos'].T).T
            assert (frac > -1e-9).all() and (frac < 1 + 1e-9).all()
            shift = np.linalg.solve(cell.T, (g['pos'] - f['pos']).T).T
            np.testing.assert_allclose(shift, np.rint(shift), atol=1e-7)
            np.testing.assert_allclose(g['cell'], cell, atol=1e-8)
        else:
            np.testing.assert_allclose(g['pos'], f['pos'], atol=1e-9)
            assert (g['cell'] == 0).all()  # reference: cell[~pbc] = 0
    z, pos, cell_t, batch, energy, force = D.collate(got)
    assert z.shape == (18,) and pos.shape == (18, 3) and cell_t.shape == (4, 3, 3) and energy.shape == (4,)
    assert batch.tolist() == [0] * 3 + [1] * 4 + [2] * 5 + [3] * 6 and force.dtype == torch.float32


def test_units_scale(tmp_path):
    f = [{'z': np.array([1, 8]), 'pos': np.array([[0., 0, 0], [1., 0, 0]]), 'cell': np.zeros((3, 3)), 'energy': 2.0,
          'force': np.ones((2, 3))}]
    _write_extxyz(tmp_path / 'a.xyz', f, False)
    g = D.read_extxyz(str(tmp_path / 'a.xyz'), length_unit=0.5, energy_unit=4.0)[0]
    assert g['pos'][1, 0] == 0.5 and g['energy'] == 8.0 and g['force'][0, 0] == 8.0


def test_fit_scalers_sets_energy_scale_shift():
    from newtonnet_b200.models import NewtonNet
    g, frames = _frames_from_golden()
    model = NewtonNet(output_properties=['energy', 'gradient_force'])
    D.fit_scalers(model, D.molecular_statistics(frames))
    np.testing.assert_allclose(model.scalers[0].shift.weight.detach().numpy()[:, 0], g['e_shift'].astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(model.scalers[0].scale.weight.detach().numpy()[:, 0], g['e_scale'].astype(np.float32), rtol=1e-6)
    assert model.scalers[0].shift.weight.dtype == torch.float32


def test_write_extxyz_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    cell = np.diag([9.0, 10.0, 11.0])
    frames = [{'z': np.array([8, 1, 1]), 'pos': rng.uniform(0, 9, (3, 3)), 'cell': cell, 'energy': -1.5, 'force': rng.normal(size=(3, 3))},
              {'z': np.array([6, 1]), 'pos': rng.uniform(0, 9, (2, 3)), 'cell': cell, 'energy': 2.25, 'force': rng.normal(size=(2, 3))}]
    D.write_extxyz(str(tmp_path / 'w.xyz'), frames)
    D.write_extxyz(str(tmp_path / 'w.xyz'), [{'z': np.array([1]), 'pos': np.zeros((1, 3))}], append=True)
    got = D.read_extxyz(str(tmp_path / 'w.xyz'))
    assert len(got) == 3 and got[2]['energy'] is None and got[2]['force'] is None and (got[2]['cell'] == 0).all()
    for f, g in zip(frames, got):
        assert (f['z'] == g['z']).all() and abs(f['energy'] - g['energy']) < 1e-9
        np.testing.assert_allclose(g['pos'], f['pos'], atol=1e-8)
        np.testing.assert_allclose(g['force'], f['force'], atol=1e-8)
        np.testing.assert_allclose(g['cell'], cell, atol=1e-8)
