"""TEST INFRASTRUCTURE - CPU restatement of the device-resident MD step (newtonnet_b200/csrc/md_ops.cu).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

What it pins:
  * philox4x32_10: the published Philox4x32-10 counter-based generator (Salmon et al., SC'11, Random123); checked
    against Random123's known-answer vectors in tests/test_md_host.py.
  * baoab_step: B (half kick) A (half drift) O (exact Ornstein-Uhlenbeck) A (half drift) - force - B (half kick);
    with ou_c = 1 this is the velocity-Verlet step that ASE's VelocityVerlet takes around
    MLAseCalculator.calculate (reference utils/ase_interface.py:52-81, scripts/simulate.py:21-31 use ASE's Langevin,
    a third-party scheme not present in /root/reference: the stochastic integrator's parity is statistical).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = [int(x) & MASK for x in counter]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def normal3(seed, atom, step):
    """Three standard normals for (atom, step): counter = (atom, step_lo, step_hi, 0x4D44), key = seed (lo, hi)."""
    c = philox4x32_10([atom, step & MASK, (step >> 32) & MASK, 0x4D44], [seed & MASK, (seed >> 32) & MASK])
    u = [(x + 0.5) / 4294967296.0 for x in c]
    r0, r1 = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    return np.array([r0 * np.cos(2 * np.pi * u[1]), r0 * np.sin(2 * np.pi * u[1]), r1 * np.cos(2 * np.pi * u[3])])


def baoab_half(x, v, f, inv_mass, dt, ou_c=1.0, kT=0.0, seed=0, step=0):
    """B A O A: returns (x', v') before the new force is known.  x, v, f [N,3] fp64; inv_mass [N]."""
    h = 0.5 * dt
    u = v + h * inv_mass[:, None] * f
    p = x + h * u
    if ou_c < 1.0:
        g = np.stack([normal3(seed, i, step) for i in range(len(x))])
        u = ou_c * u + np.sqrt((1.0 - ou_c * ou_c) * kT * inv_mass)[:, None] * g
    return p + h * u, u


def kick(v, f, inv_mass, dt):
    return v + 0.5 * dt * inv_mass[:, None] * f


def wrap(pos, cell):
    """Fractional coordinates into [0, 1) (rows of `cell` are lattice vectors); zero cell = unchanged."""
    if np.linalg.det(cell) == 0.0:
        return pos
    frac = pos @ np.linalg.inv(cell)
    return (frac - np.floor(frac)) @ cell
