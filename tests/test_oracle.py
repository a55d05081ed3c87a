"""Pins the CPU oracle (oracle/newtonnet_oracle.py) to the reference.

Golden sources: the reference's own MD trajectory scripts/md17_md/md.traj + shipped checkpoint, and
outputs of the unmodified reference run in the build container (tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case, load_weights
from oracle import newtonnet_oracle as O

CASES_EF = ['aspirin1', 'aspirin100', 'mols24', 'mols_edge']
CASES_PBC = ['water375', 'water81_smallL', 'water192_ortho_unwrapped', 'water_batch2', 'water192_triclinic']


def test_known_answer_md_traj():
    """All 201 frames the reference calculator wrote (fp32, CUDA) - energies bit-equal in fp32,
    forces within the reference's own fp32 noise (1.6e-5 eV/A)."""
    kat = np.load(f'{GOLDEN}/md17_kat.npz')
    w = load_weights('md17')
    nf = kat['positions'].shape[0]
    assert nf == 201
    z = np.tile(kat['numbers'], nf)
    pos = kat['positions'].reshape(-1, 3).astype(np.float32)
    batch = np.repeat(np.arange(nf), 21)
    out = O.forward(w, z, pos, np.zeros((nf, 3, 3), np.float32), batch, dtype=torch.float32)
    assert abs(out['energy'][0] - (-17591.8262)) < 2e-3          # scripts/md17_md/md.log:2
    np.testing.assert_allclose(out['energy'], kat['energy'], rtol=0, atol=2e-3)   # 1 ulp at 1.76e4 eV
    assert np.abs(out['forces'].reshape(nf, 21, 3) - kat['forces']).max() < 5e-5
    out64 = O.forward(w, z, pos, np.zeros((nf, 3, 3), np.float32), batch, dtype=torch.float64)
    assert np.abs(out64['energy'] / kat['energy'] - 1).max() < 1e-6
    assert np.abs(out64['forces'].reshape(nf, 21, 3) - kat['forces']).max() < 1e-4


@pytest.mark.parametrize('name', CASES_EF + CASES_PBC)
def test_forward_matches_reference(name):
    d, w = load_case(name)
    stress = 'stress' in list(d['props'])
    o32 = O.forward(w, d['z'], d['pos'], d['cell'], d['batch'], dtype=torch.float32, stress=stress)
    assert np.array_equal(o32['edge_index'], d['ref32_edge_index'])      # same order, bit exact
    np.testing.assert_allclose(o32['energy'], d['ref32_energy'], rtol=2e-6, atol=1e-5)
    assert np.abs(o32['forces'] - d['ref32_forces']).max() < 1e-5
    o64 = O.forward(w, d['z'], d['pos'], d['cell'], d['batch'], dtype=torch.float64, stress=stress)
    np.testing.assert_allclose(o64['energy'], d['ref64_energy'], rtol=1e-12, atol=1e-10)
    assert np.abs(o64['forces'] - d['ref64_forces']).max() < 1e-12
    if stress:
        assert np.abs(o64['stress'] - d['ref64_stress']).max() < 1e-14
        assert np.abs(o64['virial'] - d['ref64_virial']).max() < 1e-11
    if 'ref64_atom_node' in d:
        assert np.abs(o64['atom_node'] - d['ref64_atom_node']).max() < 1e-11
        assert np.abs(o64['force_node'] - d['ref64_force_node']).max() < 1e-11


@pytest.mark.parametrize('name', CASES_EF + CASES_PBC)
def test_analytic_backward_matches_reference_autograd(name):
    """Row B of SURVEY.md §8a: the hand-derived pair-symmetric reverse sweep equals autograd."""
    d, w = load_case(name)
    oa = O.forward_analytic(w, d['z'], d['pos'], d['cell'], d['batch'])
    np.testing.assert_allclose(oa['energy'], d['ref64_energy'], rtol=1e-12, atol=1e-10)
    assert np.abs(oa['forces'] - d['ref64_forces']).max() < 1e-12
    if 'ref64_stress' in d:
        assert np.abs(oa['stress'] - d['ref64_stress']).max() < 1e-14
        assert np.abs(oa['virial'] - d['ref64_virial']).max() < 1e-11


@pytest.mark.parametrize('name', ['water375', 'water81_smallL', 'water1029', 'water192_ortho_unwrapped',
                                  'water_batch2', 'mols24', 'mols_edge', 'mols256', 'aspirin100'])
def test_cell_list_restatement_is_bit_exact(name):
    d, _ = load_case(name)
    ei, disp = O.radius_graph_cell_list(d['pos'], d['cell'], d['batch'])
    assert np.array_equal(ei, d['ref32_edge_index'])
    ed, dd = O.radius_graph_dense(torch.tensor(d['pos']), torch.tensor(d['cell']), torch.tensor(d['batch']))
    assert np.array_equal(dd.numpy(), disp)


def test_edge_set_properties():
    """Symmetric edge set, one image per ordered pair even when L < 2 rc."""
    d, _ = load_case('water81_smallL')
    ei = d['ref32_edge_index']
    fw = set(map(tuple, ei.T.tolist()))
    assert len(fw) == ei.shape[1]                       # no duplicate ordered pair
    assert all((j, i) in fw for (i, j) in fw)           # symmetric
    assert all(i != j for (i, j) in fw)


def test_cutoff_envelope_factorisation():
    """1 - 55x^9 + 99x^10 - 45x^11 == (1-x)^3 * sum_k C(k+2,2) x^k, k=0..8 (used by the CUDA kernels)."""
    x = torch.linspace(0, 1, 1001, dtype=torch.float64).unsqueeze(1)
    coef = torch.tensor([(k + 1) * (k + 2) / 2 for k in range(9)], dtype=torch.float64)
    fact = (1 - x) ** 3 * (coef * x.pow(torch.arange(9))).sum(1, keepdim=True)
    assert torch.allclose(O.polynomial_cutoff(x), fact, atol=1e-12)


def test_synthetic_generators():
    z, pos, cell, batch = O.water_box(10)
    assert len(z) == 3000 and abs(cell[0, 0, 0] - 31.04) < 1e-5
    ei, _ = O.radius_graph_cell_list(pos, cell, batch)
    assert 50 < ei.shape[1] / 3000 < 58
    z, pos, cell, batch = O.molecule_batch(64)
    assert batch.max() == 63 and np.all(np.diff(batch) >= 0)


@pytest.mark.parametrize('name', ['mols24', 'water81'])
def test_training_gradients_match_reference(name):
    """Row T: loss and parameter gradients of one training step (double backward through the forces)
    against the unmodified reference in train mode (tests/golden/train_*.npz)."""
    d = dict(np.load(f'{GOLDEN}/train_{name}.npz'))
    w = load_weights('seed0')
    loss, g = O.training_gradients(w, d['z'], d['pos'], d['cell'], d['batch'], d['e_target'], d['f_target'],
                                   float(d['force_weight']))
    assert abs(loss - float(d['loss'])) < 1e-9 * abs(float(d['loss']))
    for k, v in g.items():
        assert np.abs(v - d['grad.' + k]).max() < 1e-9 * max(1.0, np.abs(d['grad.' + k]).max()), k
    # the dead layer-0 equiv_message2 gets exactly zero gradient
    assert np.abs(d['grad.interaction_layers.0.equiv_message2.0.weight']).max() == 0.0


@pytest.mark.parametrize('name', ['mols_edge', 'water81'])
def test_layer_norm_and_direct_force_match_reference(name):
    """layer_norm=True and the direct_force head against the unmodified reference."""
    d = dict(np.load(f'{GOLDEN}/wide_{name}.npz'))
    w = load_weights('seed0')
    w.update({k[6:]: v for k, v in d.items() if k.startswith('extra.')})
    o = O.forward(w, d['z'], d['pos'], d['cell'], d['batch'], direct_force_head=2)
    np.testing.assert_allclose(o['energy'], d['ref64_energy'], rtol=1e-12, atol=1e-10)
    assert np.abs(o['forces'] - d['ref64_forces']).max() < 1e-12
    assert np.abs(o['direct_force'] - d['ref64_direct_force']).max() < 1e-12


def test_hessian_matches_reference():
    d = np.load(f'{GOLDEN}/hessian_aspirin1.npz')
    w = load_weights('md17')
    h = O.hessian(w, d['z'], d['pos'], np.zeros((1, 3, 3)), np.zeros(21, dtype=np.int64))
    assert h.shape == (21, 3, 21, 3)
    assert np.abs(h - d['hessian']).max() < 1e-4        # golden positions are stored in fp32


def test_cluster_cut_reproduces_periodic_forces():
    """oracle/cluster.py, the checker of the 98k-atom box (tests/test_gpu_c4.py, bench.py's in-run parity record): forces of
    the atoms at the centre of a NON-periodic cluster of radius r_dest + 2 * n_layers * cutoff equal the forces of the full
    periodic system.  Pinned here where the full periodic oracle still runs: a one-layer model (receptive field 10 A) on
    a 1,536-atom water box (L = 24.8 A)."""
    from oracle.cluster import oracle_cluster_forces
    w = load_weights('seed0')
    w1 = {k: v for k, v in w.items() if not k.startswith('interaction_layers.') or k.startswith('interaction_layers.0.')}
    assert O.n_layers(w1) == 1
    z, pos, cell, batch = O.water_box(8)
    ei, disp = O.radius_graph_cell_list(pos, cell, batch)
    full = O.forward_analytic(w1, z, pos, cell, batch, dtype=torch.float64, edge_index=ei, disp=disp)
    for center in (0, 700, 1535):
        idx, f, info = oracle_cluster_forces(z, pos, cell, w1, center, n_layers=1)
        assert info['destination_atoms'] >= 1 and info['cluster_atoms'] < len(z)
        ref = np.asarray(full['forces'])[idx]
        assert np.abs(ref).max() > 1e-3
        # the cluster's positions are re-centred and rounded to fp32 again: displacements differ by ~1e-6 A
        assert np.abs(f - ref).max() < (2e-5 if info['oracle_dtype'] == 'float64' else 1e-4), (center, info)
    # one layer too few in the radius and the forces at the centre are wrong: the check has teeth
    idx, f, info = oracle_cluster_forces(z, pos, cell, w1, 700, n_layers=0)
    assert np.abs(f - np.asarray(full['forces'])[idx]).max() > 1e-3
