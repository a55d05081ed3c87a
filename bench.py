#!/usr/bin/env python
"""Benchmark of the NewtonNet energy+force path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

metric: energy+force atom-steps per second (BASELINE.json).  One "step" = neighbour-list rebuild + one
energy+forces evaluation of one batch of synthetic input (positions change every step).
Default workload `c2` = BASELINE.json configs[1]: ANI-1x-shaped ragged batch, 4096 molecules of 4..64
atoms per GPU; with N GPUs every rank owns its own batch (independent molecules, no data-path
collective) -> weak scaling, value = atoms of all ranks / max-over-ranks device time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'energy+force atom-steps/sec'
UNIT = 'atom-steps/s'
CPU_SAMPLE_MOLECULES = {'c1': 100, 'c2': 96, 'c3': None, 'c4': None}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--backend', default=os.environ.get('NN_GEMM_BACKEND', 'auto'), choices=['auto', 'simt', 'tc', 'ts'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor_burst=d['bf16_tflops'], tensor=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured')
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source='fallback')


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active') and not v.lower().startswith('not'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_sample(workload, seed=0):
    from newtonnet_b200 import workloads
    z, pos, cell, batch = workloads.make(workload, seed)
    n_mol = CPU_SAMPLE_MOLECULES[workload]
    if n_mol is None:
        if workload == 'c4':          # the dense O(N^2) reference search cannot run 98k atoms: C3 is the proxy
            z, pos, cell, batch = workloads.make('c3', seed)
            return (z, pos, cell, batch), 'proxy: whole c3 box (3000 atoms); the reference algorithm is O(N^2) and cannot run c4'
        return (z, pos, cell, batch), f'whole {workload} system ({len(z)} atoms)'
    n = int((batch < n_mol).sum())
    return (z[:n], pos[:n], cell[:n_mol], batch[:n]), f'first {n_mol} molecules of the {workload} batch ({n} atoms)'


def seed0_weights():
    """Default-init NewtonNet(seed 0) with randomised scale/shift as a numpy state dict."""
    import torch
    from newtonnet_b200.models import NewtonNet
    torch.manual_seed(0)
    m = NewtonNet(output_properties=['energy', 'gradient_force'])
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        m.scalers[0].scale.weight.copy_(torch.rand(119, 1, generator=g) + 0.5)
        m.scalers[0].shift.weight.copy_(torch.randn(119, 1, generator=g))
    return m


def run_oracle_timed(workload, steps, warmup, budget_s=None):
    """The reference's algorithm (oracle port: the same PyTorch-CPU op sequence incl. autograd forces)
    on all host cores, fp32."""
    import torch
    from oracle import newtonnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    (z, pos, cell, batch), sample = cpu_sample(workload)
    sd = {k: v.detach().numpy() for k, v in seed0_weights().state_dict().items()}
    stress = workload in ('c3', 'c4')
    rng = np.random.default_rng(0)
    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        p = (pos + rng.normal(0, 0.01, pos.shape)).astype(np.float32)
        t0 = time.perf_counter()
        O.forward(sd, z, p, cell, batch, dtype=torch.float32, stress=stress)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if budget_s is not None and it >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    total = float(np.sum(times))
    return dict(value=len(z) * len(times) / total, unit=UNIT, cores=cores, kind='port',
                sample=f'{sample}; {len(times)} timed evaluations, fp32, torch {torch.__version__} CPU, '
                       f'{cores} threads'), total / len(times), len(times), len(z)


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from newtonnet_b200 import workloads
    base, sec_per_step, n_timed, n_atoms = run_oracle_timed(args.workload, args.steps, args.warmup)
    line = {'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': n_timed,
            'warmup': args.warmup, 'ms_per_step': sec_per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': f'{args.workload}: {workloads.DESCRIPTION[args.workload]}',
                       'reference_sample': base['sample'], 'atoms_per_step': n_atoms},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- this repo's arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    from newtonnet_b200 import _lib as L
    from newtonnet_b200 import workloads
    from newtonnet_b200.engine import get_engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world == 1 and args.gpus > 1:       # convenience: relaunch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29511', os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # NCCL's version banner must not land on stdout (one JSON line)
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lib = L.load()
    backend = args.backend
    if backend == 'auto':
        backend = {0: 'simt', 1: 'tc', 2: 'ts'}[lib.nn_get_gemm_backend()]
    lib.nn_set_gemm_backend({'simt': 0, 'tc': 1, 'ts': 2}[backend])

    K, W = args.steps, max(args.warmup, 3)
    if args.workload == 'c4' and world > 1:
        return main_c4_decomposed(args, world, rank, local, dev, lib, backend, barrier, max_over_ranks)
    if args.workload == 'c5':
        return main_c5_training(args, world, rank, local, dev, lib, backend, barrier, max_over_ranks, sum_over_ranks)
    z_h, pos_h, cell_h, batch_h = workloads.make(args.workload, seed=rank)
    N, B = len(z_h), cell_h.shape[0]
    stress = args.workload in ('c3', 'c4')
    props = ['energy', 'gradient_force'] + (['stress'] if stress else [])
    model = seed0_weights()
    model.output_properties = props
    if stress:
        from newtonnet_b200.models.output import get_aggregator_by_string, get_output_by_string
        from newtonnet_b200.layers.scalers import get_scaler_by_string
        model.output_layers.append(get_output_by_string('stress'))
        model.scalers.append(get_scaler_by_string('stress'))
        model.aggregators.append(get_aggregator_by_string('stress'))
    model = model.to(dev)
    model.eval()
    model.return_node_features = False
    engine = get_engine(dev)
    pack = model._weight_pack(dev)

    # positions change every step (sigma = 0.01 A, like consecutive MD frames)
    rng = np.random.default_rng(100 + rank)
    steps_pos = [(pos_h + rng.normal(0, 0.01, pos_h.shape)).astype(np.float32) for _ in range(K + W)]
    z_d = torch.tensor(z_h, device=dev); cell_d = torch.tensor(cell_h, device=dev); batch_d = torch.tensor(batch_h, device=dev)
    pos_d = [torch.tensor(p, device=dev) for p in steps_pos]

    def step_device(i):
        nl = engine.neighbor_list(pos_d[i], cell_d, batch_d, pack.cutoff)
        out = engine.evaluate(nl, pack, z_d, want_forces=True, want_virial=stress)
        return nl, out

    # first call sizes the capacities (one host sync), then steady state is sync-free
    nl, out = step_device(0)
    st = nl.check()
    n_edges = st[L.ST_N_EDGES]
    for i in range(W):
        nl, out = step_device(i)
    torch.cuda.synchronize()
    assert nl.check()[L.ST_EDGE_OVERFLOW] == 0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.nn_profile_enable(1)
    lib.nn_launch_count(1)
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(W, W + K):
        nl, out = step_device(i)
    ev1.record()
    torch.cuda.synchronize(); barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = int(lib.nn_launch_count(1))
    ms = (C.c_float * 12)(); cnt = (C.c_int * 12)()
    lib.nn_profile_collect(ms, cnt, 12)
    lib.nn_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    st = nl.check()
    assert st[L.ST_EDGE_OVERFLOW] == 0, 'capacity overflow inside the timed region'
    e_last = out['energy'].double().sum().item()
    assert np.isfinite(e_last)
    dev_ms = max_over_ranks(dev_ms)
    atoms_all = sum_over_ranks(float(N))
    value = atoms_all * K / (dev_ms * 1e-3)

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory()
        z_p, cell_p, batch_p = pin(z_h), pin(cell_h), pin(batch_h)
        pos_p = [pin(p) for p in steps_pos]
        e_host = torch.empty(B, dtype=torch.float32).pin_memory()
        f_host = torch.empty(N, 3, dtype=torch.float32).pin_memory()

        def step_e2e(i):
            zt = z_p.to(dev, non_blocking=True); pt = pos_p[i].to(dev, non_blocking=True)
            ct = cell_p.to(dev, non_blocking=True); bt = batch_p.to(dev, non_blocking=True)
            o = model(zt, pt, ct, bt)
            e_host.copy_(o.energy, non_blocking=True)
            f_host.copy_(o.gradient_force, non_blocking=True)
            torch.cuda.synchronize()

        api = 'newtonnet_b200.NewtonNet.forward(z,pos,cell,batch) on pinned host inputs, results copied to host'
        if args.workload == 'c3':
            # config 3 is an ASE-calculator MD step: go through MLAseCalculator.calculate with a duck-typed Atoms
            # (ase is not installed in the image): numpy in, numpy out, wrapped positions, Voigt stress
            from newtonnet_b200.utils.ase_interface import MLAseCalculator
            calc = MLAseCalculator(model, properties=['energy', 'forces', 'stress'], device=str(dev))

            class Atoms:
                def __init__(self, pos): self.pos = pos
                def __len__(self): return N
                def copy(self): return self
                def get_atomic_numbers(self): return z_h
                def get_positions(self, wrap=False): return self.pos
                def get_cell(self): return cell_h[0].astype(np.float64)
                def get_pbc(self): return np.array([True, True, True])
            frames = [Atoms(p.astype(np.float64)) for p in steps_pos]

            def step_e2e(i):   # noqa: F811
                calc.calculate(frames[i])
                assert calc.results['forces'].shape == (N, 3)
            api = 'newtonnet_b200.utils.ase_interface.MLAseCalculator.calculate(atoms) (numpy in, numpy out; energy, forces, stress)'
        for i in range(W):
            step_e2e(i)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(W, W + K):
            step_e2e(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        e2e_s = max_over_ranks(e2e_s)
        h2d = z_h.nbytes + steps_pos[0].nbytes + cell_h.nbytes + batch_h.nbytes
        d2h = e_host.numel() * 4 + f_host.numel() * 4
        e2e = {'value': atoms_all * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_s / K * 1e3,
               'api': api}

    # ---- device-resident MD (caller side): same box, integrator state in HBM, one CUDA graph per step
    md_info = None
    if args.workload in ('c1', 'c3') and world == 1 and not args.no_e2e:
        try:
            from newtonnet_b200.md import DeviceMD, FS
            md = DeviceMD(model, z_h, steps_pos[0].astype(np.float64), cell=cell_h, batch=batch_h, temperature_K=300.0,
                          friction=1.0 / (500 * FS), timestep=0.5 * FS, check_interval=max(K, 1))
            md.run(max(W, 3))
            n_md = max(K, 20)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            md.run(n_md)                                           # ends with the status / log read (synchronises)
            md_s = time.perf_counter() - t0
            md_info = {'ms_per_step': md_s / n_md * 1e3, 'value': N * n_md / md_s, 'unit': UNIT, 'steps': n_md,
                       'kernels_per_step': md.kernels_per_step, 'graph_replays': n_md,
                       'what': 'newtonnet_b200.md.DeviceMD.run: Langevin (BAOAB) step + neighbour rebuild + energy/forces, state '
                               'resident in HBM, wall clock including the per-chunk status/log read; no stress'}
        except RuntimeError as exc:      # random-weight potential energy surfaces can collapse a box; report, do not fail the bench
            md_info = {'error': str(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel class (stage times from CUDA events inside the timed region)
    pk = peaks()
    P = n_edges // 2
    stage_ms = {L.STAGES[k]: float(ms[k]) for k in range(12)}
    stage_n = {L.STAGES[k]: int(cnt[k]) for k in range(12)}
    n_layers = pack.n_layers
    F = 128
    chained = backend == 'ts' and os.environ.get('NN_CHAIN', '1') != '0'
    # algorithmic work per launch (DESIGN.md section "kernels"): GEMM = 2*M*128*128 flop and 2*M*512 B
    alg = {
        # 32 flop per byte algorithmic (2*128*128 flop per 1 KB row) << machine balance: the GEMMs are HBM-bound
        'pair_gemm': dict(bound='hbm', per_launch=None, unit='GB/s', scale=1e-9, peak=pk['hbm']),
        'node_gemm': dict(bound='tensor', per_launch=None, unit='TFLOP/s', scale=1e-12, peak=pk['tensor']),
        'message': dict(bound='hbm', per_launch=P * (512 + 8 + 80) + N * 512, unit='GB/s', scale=1e-9, peak=pk['hbm']),
        'aggregate': dict(bound='hbm', per_launch=None, unit='GB/s', scale=1e-9, peak=pk['hbm']),
        'bwd_gather': dict(bound='hbm', per_launch=None, unit='GB/s', scale=1e-9, peak=pk['hbm']),
        'bwd_message': dict(bound='hbm', per_launch=P * (1024 + 8 + 160) + 2 * N * 512, unit='GB/s', scale=1e-9, peak=pk['hbm']),
        'bwd_aggregate': dict(bound='hbm', per_launch=None, unit='GB/s', scale=1e-9, peak=pk['hbm']),
    }
    # stages whose launches differ (first layer skips the e2 / f_j streams): use the per-step total instead
    per_step_total = {
        # per pair, two launches per MLP: layer 0 (U path only) fwd 1024 + 1536, bwd 1536 + 1024; other layers fwd 2*(1024 + 1536),
        # bwd 2*1536 + 1024 + 1536.  With the chained two-CTA kernel (gemm_chain.cu: forward MLPs and the accumulating reverse
        # MLP) the intermediate never reaches HBM: forward MLP 1536, accumulating reverse MLP 2048.
        'pair_gemm': P * ((4096.0 + (n_layers - 1) * 7680.0) if chained else (5120.0 + (n_layers - 1) * 10752.0)),
        'node_gemm': 2.0 * F * F * N * (n_layers * (2 + 3 + 3 + 2) + 4),
        'aggregate': (n_layers * (P * 2 * 512 + N * (512 * 2 + 1536) + 2 * P * 8 + P * 12) + (n_layers - 1) * (P * 512 + N * 1536 * 2)),
        'bwd_gather': (n_layers * (P * (1024 + 8 + 24) + N * 1536) + (n_layers - 1) * (P * 512 + N * 1536)),
        'bwd_aggregate': (n_layers * (P * 512 + N * 1024 + 2 * P * 8) + (n_layers - 1) * (P * 512 + N * 1536 * 2)),
    }
    table = {}
    for name, a in alg.items():
        if stage_n[name] == 0 or stage_ms[name] <= 0:
            continue
        per_launch = a['per_launch'] if a['per_launch'] is not None else per_step_total[name] * K / stage_n[name]
        avg_s = stage_ms[name] * 1e-3 / stage_n[name]
        ach = per_launch / avg_s * a['scale']
        table[name] = dict(bound=a['bound'], achieved=ach, peak=a['peak'], unit=a['unit'], frac=ach / a['peak'],
                           launches=stage_n[name], ms_total=stage_ms[name], share=stage_ms[name] / dev_ms)
    dominant = max(table, key=lambda k: table[k]['ms_total']) if table else None
    roofline = None
    if dominant:
        t = table[dominant]
        roofline = {'kernel': dominant + ('[tcgen05 3xTF32]' if backend in ('tc', 'ts') and 'gemm' in dominant else
                                          ('[fp32 SIMT]' if 'gemm' in dominant else '')),
                    'bound': t['bound'], 'achieved': t['achieved'], 'peak': t['peak'], 'unit': t['unit'],
                    'frac': t['frac'], 'traffic': None, 'peak_source': pk['source'] + (' sustained bf16' if t['bound'] == 'tensor' else ' copy'),
                    'share_of_step': t['share'], 'launches_timed': t['launches']}
        if dominant == 'pair_gemm':   # the same launches as fp32-equivalent FLOP/s (one 2*128*128 product per row)
            roofline['tflops_fp32_equivalent'] = 2.0 * P * F * F * t['launches'] / K / (t['ms_total'] * 1e-3 / K) * 1e-12
            roofline['tensor_pipe_passes'] = 3
        if dominant == 'pair_gemm' and args.workload == 'c2' and chained:
            # ncu dram__bytes_read + dram__bytes_write of every pair-level GEMM launch of one step (profiles/r1c_gemm_dram.csv):
            # 5 x chain<fwd> 2.706 GB + 3 x (MUL 2.731 + plain 1.788) + 2 x chain<bwd,add> 3.655 = 34.40 GB in 13 launches;
            # algorithmic 19,456 B x 1,802,624 pairs = 35.07 GB
            roofline['traffic'] = 34.40e9 / 13
            roofline['traffic_note'] = 'average DRAM bytes per pair-level GEMM launch (ncu, 13 launches per step); algorithmic 35.07e9 / 13'
        elif dominant == 'pair_gemm' and args.workload == 'c2':
            roofline['traffic'] = 2.734e9
            roofline['traffic_note'] = 'bytes per launch of the <NONE,MUL> variant (ncu dram__bytes_read+write); algorithmic 2.769e9'

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _, _ = run_oracle_timed(args.workload, steps=40, warmup=1, budget_s=15.0)   # ~15 s of CPU work

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {workloads.DESCRIPTION[args.workload]}', 'atoms_per_gpu': N,
                   'systems_per_gpu': B, 'directed_edges_per_gpu': n_edges, 'n_features': 128, 'n_basis': 20,
                   'n_interactions': n_layers, 'cutoff': pack.cutoff, 'heads': props,
                   'gemm_backend': {'simt': 'fp32 SIMT', 'tc': 'tcgen05 3xTF32 (A, B in smem)', 'ts': 'tcgen05 3xTF32 (A in TMEM)'}[backend] + (', chained two-CTA MLPs' if chained else ''),
                   'parallelism': f'dp{world} (independent batches, no data-path collective)',
                   'cache': 'per-step working set (pair tensors, %.1f GB) exceeds the 126 MB L2; positions change every step'
                            % (n_layers * 5 * P * 512 / 1e9)},
        'gpu_launches': launches, 'clocks': clocks, 'e2e': e2e, 'roofline': roofline, 'cpu_baseline': cpu_base,
        'stages': table,
    }
    if md_info:
        line['md_device_resident'] = md_info
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_c4_decomposed(args, world, rank, local, dev, lib, backend, barrier, max_over_ranks):
    """config 4: ONE 98,304-atom periodic box split into bricks (spatial domain decomposition, halo
    exchange over NCCL/NVLink) - strong scaling: the total work is fixed, value = atoms * steps / time."""
    import torch
    import torch.distributed as dist
    from newtonnet_b200 import _lib as L
    from newtonnet_b200 import workloads
    from newtonnet_b200.distributed import DomainDecomposition
    K, W = args.steps, max(args.warmup, 3)
    z_h, pos_h, cell_h, batch_h = workloads.make('c4', seed=0)
    N = len(z_h)
    model = seed0_weights().to(dev)
    model.eval()
    dd = DomainDecomposition(model, transport=os.environ.get('NN_DD_TRANSPORT', 'p2p'))
    rng = np.random.default_rng(100)
    steps_pos = [(pos_h + rng.normal(0, 0.01, pos_h.shape)).astype(np.float32) for _ in range(K + W)]
    z_d = torch.tensor(z_h, device=dev); cell_d = torch.tensor(cell_h, device=dev)
    pos_d = [torch.tensor(p, device=dev) for p in steps_pos]
    for i in range(W):
        out = dd(z_d, pos_d[i], cell_d)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.nn_launch_count(1)
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for i in range(W, W + K):
        out = dd(z_d, pos_d[i], cell_d)
    ev1.record()
    torch.cuda.synchronize(); barrier()
    wall = time.perf_counter() - t0
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = int(lib.nn_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        value = N * K / (dev_ms * 1e-3)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'c4: ' + workloads.DESCRIPTION['c4'], 'atoms_total': N,
                           'parallelism': f'spatial domain decomposition, {dd.plan.grid} bricks, halo exchange of ghost '
                                          f'feature rows (6 exchanges per step, transport {dd.transport}: '
                                          f'{"pack kernel storing into peer memory over NVLink" if dd.transport == "p2p" else "pack kernel + NCCL all_to_all"}) '
                                          f'+ all-reduce of forces/energy/virial',
                           'owned_atoms_rank0': dd.plan.n_owned, 'ghost_atoms_rank0': dd.plan.n_ghost, 'plans_built': dd.n_plans,
                           'plan_skin_A': dd.skin,
                           'gemm_backend': {'simt': 'fp32 SIMT', 'tc': 'tcgen05 3xTF32 (A, B in smem)', 'ts': 'tcgen05 3xTF32 (A in TMEM)'}[backend],
                           'note': 'every step rebuilds the neighbour list; the brick/ghost plan (host side) is reused while no '
                                   'atom moved more than skin/2; positions resident on every rank, results complete '
                                   'on every rank'},
                'gpu_launches': launches, 'clocks': clocks,
                'e2e': {'value': N * K / max_wall(wall, max_over_ranks), 'unit': UNIT, 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0, 'note': 'wall clock around the same loop (DomainDecomposition.__call__)'},
                'roofline': None, 'cpu_baseline': None}
        print(json.dumps(line), flush=True)
    else:
        max_wall(wall, max_over_ranks)
    dist.destroy_process_group()
    return 0


def main_c5_training(args, world, rank, local, dev, lib, backend, barrier, max_over_ranks, sum_over_ranks):
    """config 5: training step (energy + force loss, double backward, Adam) on MD17-shaped synthetic data,
    100 molecules x 21 atoms per GPU, data-parallel gradient all-reduce (one flat 1.6 MB bucket)."""
    import torch
    import torch.distributed as dist
    from newtonnet_b200 import workloads
    from newtonnet_b200.train import training_step
    K, W = args.steps, max(args.warmup, 3)
    z, pos, cell, batch = workloads.make('c1', seed=rank)
    N = len(z)
    rng = np.random.default_rng(7 + rank)
    t = lambda a: torch.tensor(a, device=dev)
    e_t, f_t = t(rng.standard_normal(cell.shape[0]).astype(np.float32)), t(rng.standard_normal(pos.shape).astype(np.float32))
    model = seed0_weights().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    args_t = (t(z), t(pos), t(cell), t(batch), e_t, f_t)
    # NN_TRAIN_GRAPH=1: forward + double backward replayed as one CUDA graph (static, padded edge list).  Measured on c5:
    # 18.2 ms vs 18.5 ms eager - the step is bound by ~1,400 small kernels on the GPU, not by the host - so eager is the default
    graphed = os.environ.get('NN_TRAIN_GRAPH', '0') == '1'
    if graphed:
        from newtonnet_b200.train import GraphedTrainingStep
        step = GraphedTrainingStep(model, opt, *args_t)
        training_step = lambda model, opt, *a: step(*a)     # noqa: E731,F811
    rng_p = np.random.default_rng(11 + rank)
    pos_steps = [t((pos + rng_p.normal(0, 0.01, pos.shape)).astype(np.float32)) for _ in range(K + W)]
    for i in range(W):
        loss = training_step(model, opt, args_t[0], pos_steps[i], *args_t[2:])
    lib.nn_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        loss = training_step(model, opt, args_t[0], pos_steps[W + i], *args_t[2:])
    ev1.record()
    torch.cuda.synchronize(); barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    atoms_all = sum_over_ranks(float(N))
    launches = int(lib.nn_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        v = atoms_all * K / (dev_ms * 1e-3)
        print(json.dumps({'metric': 'training ' + METRIC, 'value': v, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                          'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': 'c5: training step, 100 x 21 atoms per GPU, loss MSE(E) + 50 MSE(F), '
                                                 'double backward, clip 1.0, Adam 1e-3; new positions every step; '
                                                 + ('forward + backward replayed as one CUDA graph (GraphedTrainingStep)' if graphed
                                                    else 'eager autograd (training_step)'), 'atoms_per_gpu': N,
                                     'parallelism': f'dp{world}, one all-reduce of a flat 401,155-float gradient bucket',
                                     'final_loss': float(loss)},
                          'gpu_launches': launches, 'clocks': clocks,
                          'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                                  'note': 'training data resident on the device'},
                          'roofline': None, 'cpu_baseline': None}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def max_wall(wall, max_over_ranks):
    return max_over_ranks(wall)


if __name__ == '__main__':
    # stdout carries exactly one JSON line: anything a library prints there (e.g. NCCL's version banner)
    # is rerouted to stderr, the result line goes to the saved descriptor
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = _REAL_STDOUT
    a = parse()
    sys.exit(main_reference(a) if a.impl == 'reference' else main_b200(a))
