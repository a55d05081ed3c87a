"""CPU checks of the MD oracle (oracle/md_oracle.py): Philox known answers, integrator identities."""
import numpy as np

from oracle import md_oracle as M


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert M.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert M.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert M.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_are_standard():
    g = np.stack([M.normal3(7, i, 3) for i in range(4000)])
    assert abs(g.mean()) < 0.03 and abs(g.std() - 1.0) < 0.03
    assert abs(np.corrcoef(g[:, 0], g[:, 1])[0, 1]) < 0.05
    assert not np.allclose(M.normal3(7, 0, 3), M.normal3(7, 0, 4))


def test_velocity_verlet_conserves_harmonic_energy():
    k, m, dt = 2.0, 3.0, 0.01
    x, v = np.array([[1.0, 0.0, 0.0]]), np.zeros((1, 3))
    im = np.array([1.0 / m])
    f = -k * x
    e0 = 0.5 * k * (x ** 2).sum()
    for _ in range(2000):
        x, v = M.baoab_half(x, v, f, im, dt)
        f = -k * x
        v = M.kick(v, f, im, dt)
    e = 0.5 * k * (x ** 2).sum() + 0.5 * m * (v ** 2).sum()
    assert abs(e - e0) / e0 < 1e-4


def test_wrap():
    cell = np.array([[4.0, 0, 0], [1.0, 5.0, 0], [0, 0, 6.0]])
    p = np.array([[-0.5, 7.0, 13.0], [3.9, 0.1, -0.1]])
    w = M.wrap(p, cell)
    frac = w @ np.linalg.inv(cell)
    assert (frac >= 0).all() and (frac < 1).all()
    s = (w - p) @ np.linalg.inv(cell)
    assert np.allclose(s, np.rint(s))
    assert (M.wrap(p, np.zeros((3, 3))) == p).all()


def test_md_units_and_masses():
    from newtonnet_b200 import md
    # ASE units (CODATA 2014): 1 fs = 1e-5 sqrt(e / amu) Angstrom sqrt(amu / eV); kB in eV / K
    assert abs(md.FS - 0.09822694788464063) < 1e-15 and abs(md.KB - 8.6173303e-5) < 1e-12
    assert abs(md.ATOMIC_MASSES[1] - 1.008) < 1e-9 and abs(md.ATOMIC_MASSES[6] - 12.011) < 1e-9 and abs(md.ATOMIC_MASSES[8] - 15.999) < 1e-9
    assert len(md.ATOMIC_MASSES) == 37 and md.ATOMIC_MASSES[0] == 0.0      # H..Kr; heavier elements need explicit masses
    # equipartition: velocities drawn as sqrt(kT / m) N(0,1) give <m v^2> = kT per degree of freedom
    rng = np.random.default_rng(0)
    m = np.array([1.008, 12.011, 15.999] * 4000)
    v = rng.normal(size=(len(m), 3)) * np.sqrt(md.KB * 300.0 / m)[:, None]
    t = (m[:, None] * v ** 2).sum() / (3 * len(m) * md.KB)
    assert abs(t - 300.0) < 5.0
