"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
The decomposed evaluation on k ranks must equal the single-GPU evaluation of the same box, and a sharded
molecule batch must equal the unsharded one (multi-GPU without a cluster)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_weights

pytestmark = pytest.mark.gpu
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model(dev, props):
    from newtonnet_b200.compat import model_from_state_dict
    w = load_weights('seed0')
    m = model_from_state_dict({k: torch.tensor(v) for k, v in w.items()}, output_properties=props).to(dev)
    m.eval()
    return m


def _dd_worker(rank, world, port, nside, out_path, transport='p2p'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from newtonnet_b200.distributed import DomainDecomposition
        from newtonnet_b200 import workloads
        z, pos, cell, batch = workloads.water_box(nside, seed=3)
        t = lambda a: torch.tensor(a, device=dev)
        model = _model(dev, ['energy', 'gradient_force', 'stress', 'virial'])
        dd = DomainDecomposition(model, transport=transport)
        out = dd(t(z), t(pos), t(cell))                  # eager first step (plan, capacities)
        out2 = dd(t(z), t(pos), t(cell))                 # captured as one CUDA graph and replayed (peer transport)
        out3 = dd(t(z), t(pos), t(cell))                 # replay
        assert torch.equal(out.gradient_force, out2.gradient_force) and torch.equal(out.gradient_force, out3.gradient_force)
        assert torch.equal(out.energy, out3.energy) and torch.equal(out.stress, out3.stress)
        extra = {}
        if transport == 'p2p':
            # no-sync steps on new positions, checked afterwards; then a large move (> skin / 2) forces a replan
            rng = np.random.default_rng(5)
            pos_b = (pos + rng.normal(0, 0.02, pos.shape)).astype(np.float32)
            ob = dd(t(z), t(pos_b), t(cell), sync=False)
            st = dd.check()
            assert dd.n_plans == 1 and st[0] == 0
            pos_c = pos_b.copy()
            pos_c[::97] += np.float32(0.7)
            oc = dd(t(z), t(pos_c), t(cell))
            assert dd.n_plans == 2
            od = dd(t(z), t(pos_c), t(cell))             # graph of the new plan
            assert torch.equal(oc.gradient_force, od.gradient_force)
            # the same evaluation without the second stream / without the graph gives the same bits
            dd2 = DomainDecomposition(model, transport=transport, overlap=False, use_cuda_graph=False)
            oe = dd2(t(z), t(pos_c), t(cell))
            assert torch.equal(oe.gradient_force, oc.gradient_force) and torch.equal(oe.energy, oc.energy)
            dd2.close()
            extra = dict(fb=ob.gradient_force.cpu().numpy(), eb=ob.energy.cpu().numpy(), fc=oc.gradient_force.cpu().numpy(),
                         ec=oc.energy.cpu().numpy(), pos_b=pos_b, pos_c=pos_c)
        if rank == 0:
            ref = model(t(z), t(pos), t(cell), t(batch))
            for key, p_ in (('b', extra.get('pos_b')), ('c', extra.get('pos_c'))):
                if p_ is not None:
                    r = model(t(z), t(p_), t(cell), t(batch))
                    extra['rf' + key] = r.gradient_force.cpu().numpy(); extra['re' + key] = r.energy.cpu().numpy()
            np.savez(out_path, e=out.energy.cpu().numpy(), f=out.gradient_force.cpu().numpy(),
                     s=out.stress.cpu().numpy(), v=out.virial.cpu().numpy(), re=ref.energy.cpu().numpy(),
                     rf=ref.gradient_force.cpu().numpy(), rs=ref.stress.cpu().numpy(), rv=ref.virial.cpu().numpy(),
                     n_owned=out.n_owned, n_ghost=out.n_ghost, z=z, pos=pos, cell=cell, batch=batch, **extra)
        dist.barrier(device_ids=[rank])
        dd.close()
    finally:
        dist.destroy_process_group()


@needs2
@pytest.mark.parametrize('transport', ['p2p', 'nccl'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_domain_decomposition_matches_single_gpu(world, transport, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    path = str(tmp_path / 'dd.npz')
    mp.spawn(_dd_worker, args=(world, _free_port(), 8, path, transport), nprocs=world, join=True)
    d = np.load(path)
    assert abs(d['e'][0] - d['re'][0]) <= 1e-5 * abs(d['re'][0])
    assert np.abs(d['f'] - d['rf']).max() < 2e-5          # same fp32 kernels, different summation split
    assert np.abs(d['s'] - d['rs']).max() < 1e-4 * np.abs(d['rs']).max()
    assert np.abs(d['v'] - d['rv']).max() < 1e-4 * np.abs(d['rv']).max()
    assert d['n_ghost'] > 0
    if transport == 'p2p':
        for key in 'bc':
            assert np.abs(d['f' + key] - d['rf' + key]).max() < 2e-5
            assert abs(d['e' + key][0] - d['re' + key][0]) <= 1e-5 * abs(d['re' + key][0])
    # ... and the oracle (fp64, reference algorithm) on the same box: the north-star tolerances
    from oracle import newtonnet_oracle as O
    w = load_weights('seed0')
    ref = O.forward(w, d['z'], d['pos'], d['cell'], d['batch'], dtype=torch.float64, stress=True)
    assert abs(d['e'][0] - ref['energy'][0]) <= 1e-5 * abs(ref['energy'][0])
    assert np.abs(d['f'] - ref['forces']).max() < 1e-4


def _dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from newtonnet_b200.distributed import shard_batch
        from newtonnet_b200 import workloads
        z, pos, cell, batch = workloads.molecule_batch(300, seed=9)
        zr, pr, cr, br, sl = shard_batch(z, pos, cell, batch, rank, world)
        t = lambda a: torch.tensor(a, device=dev)
        model = _model(dev, ['energy', 'gradient_force'])
        out = model(t(zr), t(pr), t(cr), t(br))
        np.savez(os.path.join(out_dir, f'r{rank}.npz'), e=out.energy.cpu().numpy(), f=out.gradient_force.cpu().numpy(),
                 s0=sl.start, s1=sl.stop)
        if rank == 0:
            ref = model(t(z), t(pos), t(cell), t(batch))
            np.savez(os.path.join(out_dir, 'ref.npz'), e=ref.energy.cpu().numpy(), f=ref.gradient_force.cpu().numpy())
        dist.barrier(device_ids=[rank])
    finally:
        dist.destroy_process_group()


@needs2
def test_sharded_molecule_batch_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / 'ref.npz')
    e = np.concatenate([np.load(tmp_path / f'r{r}.npz')['e'] for r in range(world)])
    f = np.concatenate([np.load(tmp_path / f'r{r}.npz')['f'] for r in range(world)])
    assert np.array_equal(e, ref['e']) and np.array_equal(f, ref['f'])     # independent systems: bit identical


def _train_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from newtonnet_b200.distributed import shard_batch
        from newtonnet_b200.train import allreduce_gradients
        from newtonnet_b200 import workloads
        z, pos, cell, batch = workloads.molecule_batch(16, seed=4)
        rng = np.random.default_rng(1)
        e_t, f_t = rng.standard_normal(16).astype(np.float32), rng.standard_normal(pos.shape).astype(np.float32)
        t = lambda a: torch.tensor(a, device=dev)

        def grads_of(r):
            zr, pr, cr, br, sl = shard_batch(z, pos, cell, batch, r, world)
            a0 = int((batch < sl.start).sum()); a1 = a0 + len(zr)
            model = _model(dev, ['energy', 'gradient_force'])
            model.train()
            out = model(t(zr), t(pr).requires_grad_(True), t(cr), t(br))
            loss = torch.nn.functional.mse_loss(out.energy, t(e_t[sl])) + \
                50.0 * torch.nn.functional.mse_loss(out.gradient_force, t(f_t[a0:a1]))
            loss.backward()
            return model

        mine = grads_of(rank)
        allreduce_gradients(mine.parameters())
        every = [grads_of(r) for r in range(world)]
        for (k, p), *others in zip(mine.named_parameters(), *[m.named_parameters() for m in every]):
            if not p.requires_grad:
                continue
            want = sum((q.grad if q.grad is not None else torch.zeros_like(q)) for _, q in others) / world
            assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6 * max(1.0, float(want.abs().max()))), k
        dist.barrier(device_ids=[rank])
    finally:
        dist.destroy_process_group()


def _graphed_dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from newtonnet_b200.train import GraphedTrainingStep
        from oracle import newtonnet_oracle as O
        z, pos, cell, batch = O.water_box(4)
        rng = np.random.default_rng(rank)
        pos = (pos + rng.normal(0, 0.02, pos.shape)).astype(np.float32)
        t = lambda a, dt=None: torch.tensor(a, device=dev, dtype=dt)
        args = (t(z), t(pos), t(cell), t(batch), t(rng.standard_normal(1), torch.float32), t(rng.standard_normal(pos.shape), torch.float32))
        dense = (args[0], t((pos * 0.7).astype(np.float32)), t((cell * 0.7).astype(np.float32))) + args[3:]
        model = _model(dev, ['energy', 'gradient_force'])
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
        step = GraphedTrainingStep(model, opt, *args)
        assert step._device_skip()
        # the second batch outgrows the edge capacity on rank 1 only: every rank must skip it on the device, then redo it
        for a in (args, dense if rank == 1 else args, args):
            step(*a)
        step.settle()
        assert step.recaptures == (1 if rank == 1 else 0), (rank, step.recaptures)
        assert int(opt.state[next(iter(model.parameters()))]['step']) == 3
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                    # data parallel: identical parameters on every rank
        assert bool(torch.isfinite(flat).all())
        dist.barrier(device_ids=[rank])
    finally:
        dist.destroy_process_group()


@needs2
def test_graphed_data_parallel_step_with_overflow_on_one_rank(tmp_path):
    mp.spawn(_graphed_dp_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


@needs2
def test_data_parallel_gradient_allreduce(tmp_path):
    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
