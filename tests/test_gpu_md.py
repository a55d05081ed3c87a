"""GPU tests of the device-resident MD driver (newtonnet_b200/md.py, csrc/md_ops.cu) against the CPU oracle
(oracle/md_oracle.py) and against a host fp64 velocity Verlet driven through the public model API."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case, load_weights

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _model(weights='md17'):
    from newtonnet_b200.compat import model_from_state_dict
    m = model_from_state_dict({k: torch.tensor(v) for k, v in load_weights(weights).items()},
                              output_properties=['energy', 'gradient_force']).to(DEV)
    m.eval()
    return m


def _aspirins(n_sys):
    kat = np.load(f'{GOLDEN}/md17_kat.npz')
    z = np.tile(kat['numbers'], n_sys)
    pos = np.concatenate([kat['positions'][(7 * k) % len(kat['positions'])] for k in range(n_sys)])
    batch = np.repeat(np.arange(n_sys), 21)
    return z, pos, batch


def test_advance_kernel_matches_oracle():
    from newtonnet_b200 import _lib as L
    from oracle import md_oracle as M
    lib = L.load()
    rng = np.random.default_rng(0)
    N = 37
    cell = np.array([[7.0, 0, 0], [1.5, 8.0, 0], [0.5, -1.0, 9.0]])
    x, v, f = rng.uniform(-10, 20, (N, 3)), rng.normal(size=(N, 3)), rng.normal(size=(N, 3)).astype(np.float32)
    im = 1.0 / rng.uniform(1, 16, N)
    dt, c, kT, seed, step = 0.05, 0.9, 0.0259, 1234567890123, 5_000_000_017
    t = lambda a, d: torch.tensor(a, dtype=d, device=DEV)
    X, V, F, IM = t(x, torch.float64), t(v, torch.float64), t(f, torch.float32), t(im, torch.float64)
    CELL, BATCH = t(cell[None], torch.float32), torch.zeros(N, dtype=torch.int64, device=DEV)
    P, CTR = torch.empty(N, 3, device=DEV), torch.tensor([step], dtype=torch.int64, device=DEV)
    L.check(lib.nn_md_advance(N, X.data_ptr(), V.data_ptr(), F.data_ptr(), IM.data_ptr(), CELL.data_ptr(), BATCH.data_ptr(),
                              P.data_ptr(), dt, c, kT, seed, CTR.data_ptr(), torch.cuda.current_stream().cuda_stream), 'adv')
    xo, vo = M.baoab_half(x, v, f.astype(np.float64), im, dt, c, kT, seed, step)
    np.testing.assert_allclose(V.cpu().numpy(), vo, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(X.cpu().numpy(), xo, rtol=1e-11, atol=1e-12)
    w = M.wrap(xo, cell.astype(np.float32).astype(np.float64))
    d = (P.cpu().numpy().astype(np.float64) - w) @ np.linalg.inv(cell)
    d -= np.rint(d)                                   # an atom within rounding of a face may land on either side
    assert np.abs(d @ cell).max() < 5e-6


def test_nve_matches_host_velocity_verlet():
    from newtonnet_b200.md import DeviceMD, FS, ATOMIC_MASSES
    from oracle import md_oracle as M
    model = _model()
    z, pos, batch = _aspirins(3)
    rng = np.random.default_rng(1)
    vel = rng.normal(size=pos.shape) * 0.05
    dt, steps = 0.5 * FS, 25
    md = DeviceMD(model, z, pos, batch=batch, velocities=vel, timestep=dt)
    out = md.run(steps)
    # host loop: fp64 integrator, forces through the public forward each step (what ASE + the calculator do)
    zt, bt = torch.tensor(z, device=DEV), torch.tensor(batch, device=DEV)
    cell = torch.zeros(3, 3, 3, device=DEV)
    im = 1.0 / ATOMIC_MASSES[z]
    def force(p):
        r = model(zt, torch.tensor(p, dtype=torch.float32, device=DEV), cell, bt)
        return r.gradient_force.double().cpu().numpy(), r.energy.double().cpu().numpy()
    x, v = pos.copy(), vel.copy()
    f, _ = force(x)
    pe = []
    for _ in range(steps):
        x, v = M.baoab_half(x, v, f, im, dt)
        f, e = force(x)
        v = M.kick(v, f, im, dt)
        pe.append(e)
    assert np.abs(md.positions - x).max() < 2e-6
    assert np.abs(md.velocities - v).max() < 2e-5
    np.testing.assert_allclose(out['energy'], np.array(pe), rtol=1e-6)
    ke = np.array([0.5 * (ATOMIC_MASSES[z] * (v ** 2).sum(-1))[batch == b].sum() for b in range(3)])
    np.testing.assert_allclose(out['kinetic'][-1], ke, rtol=1e-5)
    assert md.step == steps and md.graph_launches == steps
    more = md.run(12, trajectory_interval=5)          # snapshots after steps 5 and 10 of this call
    assert more['positions'].shape == (2, len(z), 3) and more['energy'].shape == (12, 3)
    assert np.abs(more['positions'][1] - md.positions).max() > 0 and md.step == steps + 12


def test_nve_energy_conservation():
    from newtonnet_b200.md import DeviceMD, FS
    z, pos, batch = _aspirins(8)
    md = DeviceMD(_model(), z, pos, batch=batch, temperature_K=300.0, timestep=0.5 * FS, seed=3)   # MB velocities, no friction
    out = md.run(400)
    total = out['energy'] + out['kinetic']
    drift = np.abs(total - total[0]).max(axis=0)
    assert (drift < 0.02 * out['kinetic'].mean(axis=0)).all(), drift      # << kinetic energy (0.8 eV per molecule)


def test_langevin_thermostat_reaches_temperature():
    from newtonnet_b200.md import DeviceMD, FS
    z, pos, batch = _aspirins(192)
    md = DeviceMD(_model(), z, pos, batch=batch, temperature_K=300.0, friction=0.02 / FS, timestep=0.5 * FS, seed=11,
                  check_interval=200)
    md.run(500)
    temps = []
    for _ in range(10):
        md.run(20)
        temps.append(md.temperature().mean())
    assert abs(np.mean(temps) - 300.0) < 15.0, temps


def test_overflow_rollback_reproduces_trajectory():
    from newtonnet_b200.md import DeviceMD, FS
    d, w = load_case('water375')
    from newtonnet_b200.compat import model_from_state_dict
    model = model_from_state_dict({k: torch.tensor(v) for k, v in w.items()}, output_properties=['energy', 'gradient_force']).to(DEV)
    model.eval()
    kw = dict(cell=d['cell'], temperature_K=300.0, friction=0.01 / FS, timestep=0.5 * FS, seed=5, check_interval=8)
    a = DeviceMD(model, d['z'], d['pos'], **kw)
    ra = a.run(16)
    b = DeviceMD(model, d['z'], d['pos'], **kw)
    b.sticky[0] = int(b.nl.cap_edges * 1.3)            # pretend the first chunk overflowed: roll back, regrow, redo
    cap0 = b.nl.cap_edges
    rb = b.run(16)
    assert b.nl.cap_edges > cap0 and b.step == 16
    np.testing.assert_array_equal(a.positions, b.positions)
    np.testing.assert_array_equal(ra['energy'], rb['energy'])
    # periodic: the model sees wrapped coordinates
    frac = b.pos_model.double().cpu().numpy() @ np.linalg.inv(d['cell'][0].astype(np.float64))
    assert frac.min() > -1e-6 and frac.max() < 1 + 1e-6


def test_from_atoms_and_update():
    from newtonnet_b200.md import DeviceMD, FS
    from newtonnet_b200.utils.ase_interface import MLAseCalculator
    from test_gpu_parity import FakeAtoms
    kat = np.load(f'{GOLDEN}/md17_kat.npz')
    atoms = FakeAtoms(kat['numbers'], kat['positions'][0])
    moved = {}
    atoms.set_positions = lambda p: moved.setdefault('p', p)
    calc = MLAseCalculator(_model(), properties=['energy', 'forces'], device=DEV)
    md = DeviceMD.from_atoms(atoms, calc, timestep=0.5 * FS, temperature_K=300.0, friction=1 / (500 * FS))
    out = md.run(10)
    assert out['energy'].shape == (10, 1) and np.isfinite(out['energy']).all()
    calc.calculate(FakeAtoms(kat['numbers'], md.positions))
    assert abs(float(calc.results['energy']) - out['energy'][-1, 0]) < 1e-5 * abs(out['energy'][-1, 0])
    assert np.abs(calc.results['forces'] - md.forces).max() < 1e-5
    md.update_atoms(atoms)
    assert moved['p'].shape == (21, 3)
