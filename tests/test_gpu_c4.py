"""Parity at config-4 size (98,304-atom periodic water box) on ONE GPU - SURVEY.md section 8c: the dense reference
search cannot run this box, so the edge set is compared with the oracle's cell-list restatement (validated against the
dense search at c3 size in test_oracle.py) and the forces with the oracle run on a cluster cut around a chunk of
destination atoms (oracle/cluster.py).  Multi-rank runs are compared with this single-GPU result in test_gpu_multi.py
and, at full size, inside bench.py (parity record of the decomposed run)."""
import numpy as np
import pytest
import torch

from conftest import load_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c4():
    from newtonnet_b200 import workloads
    from newtonnet_b200.compat import model_from_state_dict
    z, pos, cell, batch = workloads.make('c4', seed=0)
    dev = torch.device('cuda:0')
    w = load_weights('seed0')
    model = model_from_state_dict({k: torch.tensor(v) for k, v in w.items()},
                                  output_properties=['energy', 'gradient_force', 'stress']).to(dev)
    model.eval()
    model.return_node_features = False
    t = lambda a: torch.tensor(a, device=dev)
    out = model(t(z), t(pos), t(cell), t(batch))
    return dict(z=z, pos=pos, cell=cell, batch=batch, w=w, out=out)


def test_c4_edge_set_bit_exact(c4):
    from oracle import newtonnet_oracle as O
    ei, disp = O.radius_graph_cell_list(c4['pos'], c4['cell'], c4['batch'])
    got = c4['out'].edge_index.cpu().numpy()
    assert got.shape == ei.shape
    assert np.array_equal(got, ei)                      # same edges in the reference's order (i-major, j ascending)
    nl = c4['out'].neighbor_list
    # displacements of the forward pairs, bit for bit
    P = ei.shape[1] // 2
    fwd = ei[0] < ei[1]
    assert np.array_equal(nl.pair_disp[:P].cpu().numpy(), disp[fwd])


def test_c4_forces_vs_oracle_cluster(c4):
    from oracle.cluster import oracle_cluster_forces
    pos, cell = c4['pos'], c4['cell']
    f = c4['out'].gradient_force.cpu().numpy().astype(np.float64)
    assert np.isfinite(f).all() and np.isfinite(c4['out'].stress.cpu().numpy()).all()
    center = int(np.argmin(((pos - 0.5 * cell[0, 0, 0]) ** 2).sum(1)))
    idx, f_or, info = oracle_cluster_forces(c4['z'], pos, cell, c4['w'], center)
    assert info['destination_atoms'] >= 1
    assert np.abs(f[idx] - f_or).max() < 1e-4           # north-star force tolerance, eV/A
