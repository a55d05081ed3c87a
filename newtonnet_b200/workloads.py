"""Synthetic inputs of the shapes BASELINE.json names; numpy only.

c1: MD17-aspirin-shaped batches (21 atoms x 100)         c2: ANI-1x-shaped ragged batch (<=64 atoms x 4096)
c3: periodic water box, 3,000 atoms (nside 10)           c4: periodic water box, 98,304 atoms (nside 32)
"""
import math

import numpy as np


def water_box(nside, seed=0):
    """O at (idx+0.5)*3.104 A + U(-0.2,0.2), two H at 0.9572 A / 104.52 deg under a random rotation,
    wrapped into the cubic cell L = nside*3.104 A.  Returns z [N] i64, pos [N,3] f32, cell [1,3,3] f32,
    batch [N] i64."""
    rng = np.random.default_rng(seed)
    a = 3.104
    idx = np.stack(np.meshgrid(*[np.arange(nside)] * 3, indexing='ij'), -1).reshape(-1, 3).astype(np.float64)
    n = idx.shape[0]
    O = (idx + 0.5) * a + (rng.random((n, 3)) - 0.5) * 0.4
    Q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    ang = math.radians(104.52)
    H1 = O + Q @ np.array([0.9572, 0.0, 0.0])
    H2 = O + Q @ np.array([0.9572 * math.cos(ang), 0.9572 * math.sin(ang), 0.0])
    L = nside * a
    pos = (np.stack([O, H1, H2], 1).reshape(-1, 3) % L).astype(np.float32)
    pos = np.where(pos >= np.float32(L), np.float32(0), pos)
    z = np.tile(np.array([8, 1, 1], dtype=np.int64), n)
    cell = (np.eye(3) * L).astype(np.float32)[None]
    return z, pos, cell, np.zeros(3 * n, dtype=np.int64)


def molecule_batch(n_mol, seed=1, lo=4, hi=65, sizes=None):
    """n distinct sites of a (ceil(n^(1/3))+1)^3 lattice, spacing 1.4 A, jitter +-0.25 A, z in {1,6,7,8},
    zero cells.  Returns z, pos, cell [n_mol,3,3], batch."""
    rng = np.random.default_rng(seed)
    if sizes is None:
        sizes = rng.integers(lo, hi, n_mol)
    zs, ps, bs = [], [], []
    for b, n in enumerate(sizes):
        n = int(n)
        m = int(math.ceil(n ** (1 / 3))) + 1
        sites = rng.permutation(m ** 3)[:n]
        ijk = np.stack([sites // (m * m), (sites // m) % m, sites % m], 1).astype(np.float64)
        ps.append(ijk * 1.4 + (rng.random((n, 3)) - 0.5) * 0.5)
        zs.append(np.array([1, 6, 7, 8])[rng.integers(0, 4, n)])
        bs.append(np.full(n, b, dtype=np.int64))
    return (np.concatenate(zs).astype(np.int64), np.concatenate(ps).astype(np.float32),
            np.zeros((len(sizes), 3, 3), dtype=np.float32), np.concatenate(bs))


def aspirin_like(n_mol=100, seed=2):
    """21-atom molecules with the aspirin composition C9H8O4 on jittered lattice sites (config 1 shape)."""
    z1 = np.array([6] * 9 + [8] * 4 + [1] * 8, dtype=np.int64)
    z, pos, cell, batch = molecule_batch(n_mol, seed=seed, sizes=[21] * n_mol)
    return np.tile(z1, n_mol), pos, cell, batch


def make(workload, seed=0):
    if workload == 'c1':
        return aspirin_like(100, seed=2 + seed)
    if workload == 'c2':
        return molecule_batch(4096, seed=1 + seed)
    if workload == 'c3':
        return water_box(10, seed=seed)
    if workload == 'c4':
        return water_box(32, seed=seed)
    raise ValueError(f'unknown workload {workload}')


DESCRIPTION = {
    'c1': 'MD17-aspirin-shaped batch, 21 atoms x 100 molecules, energy+forces',
    'c2': 'ANI-1x-shaped ragged batch, 4096 molecules of 4..64 atoms per GPU, energy+forces',
    'c3': 'periodic water box, 3000 atoms (L=31.04 A), neighbour rebuild + energy+forces+stress',
    'c4': 'periodic water box, 98304 atoms (L=99.33 A), neighbour rebuild + energy+forces+stress',
    'c5': 'training step on MD17-shaped batches, 21 atoms x 100 molecules per GPU',
}
