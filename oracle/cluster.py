"""Force check on a cluster cut out of a large periodic box.  TEST INFRASTRUCTURE ONLY (see newtonnet_oracle.py).

The dense reference search (layers/representations.py:72-93) is O(N^2) and cannot run a 100k-atom box, and neither can
the oracle's autograd forward.  The force on an atom depends only on atoms within 2 * n_layers * cutoff of it (the
energy of atom k sees n_layers * cutoff, and F_i collects dE_k/dr_i of every k that sees i), so the oracle is run on a
non-periodic cluster of that radius around a chunk of destination atoms ("chunked by destination
rows") and only the destination atoms' forces are compared.
"""
import numpy as np


def oracle_cluster_forces(z, pos, cell, sd, center, r_dest=1.5, n_layers=3, cutoff=5.0):
    """Forces of the destination atoms within r_dest of atom `center` from the CPU oracle (cell-list edge set +
    hand-derived reverse sweep, fp64) on a non-periodic cluster cut out of the periodic box: the force on an atom
    depends on atoms within 2 * n_layers * cutoff of it, so the cluster radius is r_dest + 2 * n_layers * cutoff + margin."""
    import os
    import torch
    from oracle import newtonnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1
    Ld = np.diag(cell[0]).astype(np.float64)
    d = pos.astype(np.float64) - pos[center].astype(np.float64)
    d -= Ld * np.rint(d / Ld)
    r = np.sqrt((d ** 2).sum(1))
    R = r_dest + 2 * n_layers * cutoff + 0.5
    assert R < 0.5 * Ld.min(), 'box too small for a cluster check'
    sel = np.nonzero(r < R)[0]
    dest = np.nonzero(r[sel] < r_dest)[0]
    cp = (pos[center].astype(np.float64) + d[sel]).astype(np.float32)
    zero = np.zeros((1, 3, 3), np.float32)
    b0 = np.zeros(len(sel), np.int64)
    ei, disp = O.radius_graph_cell_list(cp, zero, b0, cutoff)
    try:
        import psutil
        big = psutil.virtual_memory().available > 48e9
    except Exception:      # noqa: BLE001
        big = False
    dt = torch.float64 if big else torch.float32
    out = O.forward_analytic(sd, z[sel], cp, zero, b0, dtype=dt, cutoff=cutoff, edge_index=ei, disp=disp)
    f = np.asarray(out['forces'], dtype=np.float64)
    return sel[dest], f[dest], {'cluster_atoms': int(len(sel)), 'destination_atoms': int(len(dest)), 'cluster_radius_A': float(R),
                                'cluster_directed_edges': int(ei.shape[1]), 'oracle_dtype': str(dt).replace('torch.', '')}


