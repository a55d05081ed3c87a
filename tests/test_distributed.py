"""Host logic of the multi-GPU paths on CPU: brick decomposition / halo plans (numpy), the halo exchange
over a world_size-2 gloo group, molecule-batch sharding.  GPU runs of the same paths: test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from newtonnet_b200.distributed import HaloExchange, HaloPlan, shard_batch
from oracle import newtonnet_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_halo_plan_consistency(world):
    z, pos, cell, batch = O.water_box(8)           # 1,536 atoms, L = 24.8 A: 12.4 A bricks
    pos = pos + np.float32(24.832) * np.random.default_rng(0).integers(-1, 2, pos.shape)   # unwrapped input
    plans = [HaloPlan(pos, cell[0], r, world, 5.0) for r in range(world)]
    owned = np.concatenate([p.owned for p in plans])
    assert len(owned) == len(pos) and len(np.unique(owned)) == len(pos)          # a partition
    for r, p in enumerate(plans):
        assert not set(p.owned) & set(p.ghost)
        assert sum(p.recv_counts) == p.n_ghost and p.recv_counts[r] == 0
        off = 0
        for s, q in enumerate(plans):
            n = p.send_counts[s]
            assert n == q.recv_counts[r]
            sent_global = p.local_to_global[p.send_index[off:off + n]]
            start = sum(q.recv_counts[:r])
            assert np.array_equal(sent_global, q.ghost[start:start + n])        # same rows, same order
            off += n
    # every neighbour of an owned atom is present locally (edge set from the oracle's cell list)
    ei, _ = O.radius_graph_cell_list(pos, cell, batch)
    for p in plans:
        local = np.zeros(len(pos), dtype=bool)
        local[p.local_to_global] = True
        mine = np.isin(ei[0], p.owned)
        assert local[ei[1][mine]].all()


def test_halo_plan_rejects_bad_cells():
    z, pos, cell, batch = O.water_box(3)            # L = 9.3 A: two bricks would be thinner than the cutoff
    with pytest.raises(ValueError):
        HaloPlan(pos, cell[0], 0, 2, 5.0)
    tri = cell[0].copy(); tri[1, 0] = 1.0
    with pytest.raises(ValueError):
        HaloPlan(pos, tri, 0, 2, 5.0)


def _exchange_worker(rank, world, port, pos, cell):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        plan = HaloPlan(pos, cell, rank, world, 5.0)
        halo = HaloExchange(plan, 'cpu')
        for width in (128, 384):
            g = torch.from_numpy(plan.local_to_global).double()
            rows = (g[:, None] * 1000 + torch.arange(width)[None, :]).float().contiguous()
            want = rows.clone()
            rows[plan.n_owned:] = -1.0                   # ghosts unknown before the exchange
            halo.exchange(rows)
            assert torch.equal(rows, want), f'rank {rank} width {width}'
        # partial sums completed by all-reduce, as DomainDecomposition does for forces / energy
        f = torch.zeros(len(pos), 3)
        f[torch.from_numpy(plan.owned)] = 1.0
        dist.all_reduce(f)
        assert bool((f == 1.0).all())
    finally:
        dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    z, pos, cell, batch = O.water_box(6)
    mp.spawn(_exchange_worker, args=(2, _free_port(), pos, cell[0]), nprocs=2, join=True)


def test_shard_batch_balanced_partition():
    z, pos, cell, batch = O.molecule_batch(257, seed=5)
    for world in (1, 2, 4, 8):
        seen, atoms = [], []
        for r in range(world):
            zr, pr, cr, br, sl = shard_batch(z, pos, cell, batch, r, world)
            assert len(zr) == len(pr) == len(br) and cr.shape[0] == sl.stop - sl.start
            if len(br):
                assert br.min() == 0 and br.max() == cr.shape[0] - 1 and np.all(np.diff(br) >= 0)
            seen.append((sl.start, sl.stop)); atoms.append(len(zr))
        assert seen[0][0] == 0 and seen[-1][1] == 257 and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
        assert sum(atoms) == len(z) and max(atoms) - min(atoms) <= 64 * 2


def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from newtonnet_b200.train import allreduce_gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)),
                  torch.nn.Parameter(torch.randn(2, 2), requires_grad=False), torch.nn.Parameter(torch.randn(4))]
        params[0].grad = torch.full((5, 3), float(rank + 1))
        params[1].grad = torch.arange(7.0) * (rank + 1)
        # params[3] has no gradient on any rank (like the dead layer-0 equiv_message2): reduced as zeros
        flat = allreduce_gradients(params)
        assert flat.numel() == 15 + 7 + 4
        assert torch.allclose(params[0].grad, torch.full((5, 3), 1.5))
        assert torch.allclose(params[1].grad, torch.arange(7.0) * 1.5)
        assert params[2].grad is None and torch.equal(params[3].grad, torch.zeros(4))
        # condition flags ride along in the bucket and come back SUMMED (the edge-overflow flag of GraphedTrainingStep: every
        # rank must see that one of them overflowed), the gradients are still averaged
        params[0].grad = torch.full((5, 3), float(rank + 1))
        flags = torch.tensor([1.0 if rank == 1 else 0.0, 2.0])
        flat = allreduce_gradients(params, flags=flags)
        assert flat.numel() == 15 + 7 + 4
        assert torch.equal(flags, torch.tensor([1.0, 4.0]))
        assert torch.allclose(params[0].grad, torch.full((5, 3), 1.5))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    mp.spawn(_grad_worker, args=(2, _free_port()), nprocs=2, join=True)
