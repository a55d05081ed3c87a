#!/usr/bin/env python
"""Benchmark of the NewtonNet energy+force path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

metric: energy+force atom-steps per second (BASELINE.json).  One "step" = neighbour-list rebuild + one
energy+forces(+stress) evaluation of synthetic input (positions change every step).

Default workload `c4` = BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is quoted on: ONE periodic
water box of 98,304 atoms.  N = 1: the single-GPU path.  N > 1: spatial domain decomposition
(newtonnet_b200.distributed.DomainDecomposition: bricks + ghost atoms, halo exchange of feature rows as stores into
peer memory over NVLink, the whole step one CUDA graph) - STRONG scaling, value = atoms * steps / max-over-ranks time.
The line carries an in-run parity record (N-rank result against the 1-rank CUDA result and against the CPU oracle on
a cluster cut around a chunk of destination atoms) and `extra` blocks with the other configurations (c2 molecule
batches - weak scaling, no data-path collective; c3 ASE-calculator MD step; c5 training step) so that the driver's
run carries them too.  `--workload cX` makes cX the headline instead.
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
CPU_SAMPLE_MOLECULES = {'c1': 100, 'c2': 256, 'c3': None, 'c4': None}
BACKENDS = {'simt': 'fp32 SIMT', 'tc': 'tcgen05 3xTF32 (A, B in smem)', 'ts': 'tcgen05 3xTF32 (A in TMEM)'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c4', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--backend', default=os.environ.get('NN_GEMM_BACKEND', 'auto'), choices=['auto', 'simt', 'tc', 'ts'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the extra (c2 / c3 / c5) blocks')
    ap.add_argument('--no-parity', action='store_true', help='skip the in-run parity record of the decomposed c4 run')
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
                                          '-lms', '50', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
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
    if workload == 'c5':
        workload = 'c1'
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


def build_model(dev, stress):
    model = seed0_weights()
    props = ['energy', 'gradient_force'] + (['stress'] if stress else [])
    model.output_properties = props
    if stress:
        from newtonnet_b200.layers.scalers import get_scaler_by_string
        from newtonnet_b200.models.output import get_aggregator_by_string, get_output_by_string
        model.output_layers.append(get_output_by_string('stress'))
        model.scalers.append(get_scaler_by_string('stress'))
        model.aggregators.append(get_aggregator_by_string('stress'))
    model = model.to(dev)
    model.eval()
    model.return_node_features = False
    return model, props


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
            'warmup': args.warmup, 'ms_per_step': sec_per_step * 1e3, 'higher_is_better': True,
            'scaling': 'strong' if args.workload == 'c4' else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': f'{args.workload}: {workloads.DESCRIPTION[args.workload]}',
                       'reference_sample': base['sample'], 'atoms_per_step': n_atoms},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- helpers of this repo's arm
class Ctx:
    """Process-wide state of one bench run (rank, device, collectives)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from newtonnet_b200 import _lib as L
        self.args = args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # NCCL's banner must not land on stdout (one JSON line)
            dist.init_process_group('nccl', device_id=self.dev)
        self.lib = L.load()
        backend = args.backend
        if backend == 'auto':
            backend = {0: 'simt', 1: 'tc', 2: 'ts'}[self.lib.nn_get_gemm_backend()]
        self.lib.nn_set_gemm_backend({'simt': 0, 'tc': 1, 'ts': 2}[backend])
        self.backend = backend
        self.K, self.W = args.steps, max(args.warmup, 3)

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(device_ids=[self.local])

    def _reduce(self, x, op):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.MAX) if self.world > 1 else x

    def sum_over_ranks(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.SUM) if self.world > 1 else x


def traffic_record(workload):
    """DRAM bytes per launch of the dominant kernel class from this round's committed ncu capture
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py from the ncu CSV named in it)."""
    p = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if not os.path.exists(p):
        return None, 'no ncu capture committed for this round yet'
    d = json.load(open(p)).get(workload)
    if not d:
        return None, f'profiles/r2_traffic.json has no entry for {workload}'
    return d['dram_bytes_per_launch'], (f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over {d['launches']} {d['kernel_class']} "
                                        f"launches of one step; {d['source']} captured at commit {d['commit']}")


def roofline_from_stages(ctx, stage_ms, stage_n, dev_ms, K, N, P, n_layers, workload):
    """Stage table + roofline record of the dominant kernel class (stage times: CUDA events on the launching stream
    inside the timed region; algorithmic bytes: DESIGN.md section 4)."""
    pk = peaks()
    backend = ctx.backend
    chained = backend == 'ts' and os.environ.get('NN_CHAIN', '1') != '0'
    F = 128
    hbm = dict(bound='hbm', unit='GB/s', scale=1e-9, peak=pk['hbm'])
    per_step = {
        # pair-level 128x128 contractions, 32 flop per byte algorithmic (2*128*128 flop per 1 KB row) << machine balance:
        # HBM-bound.  Two launches per MLP: layer 0 (U path only) fwd 1024 + 1536, bwd 1536 + 1024; other layers fwd
        # 2*(1024 + 1536), bwd 2*1536 + 1024 + 1536.  Chained two-CTA kernel (forward MLPs and the accumulating reverse MLP):
        # the intermediate never reaches HBM: forward MLP 1536, accumulating reverse MLP 2048.
        'pair_gemm': P * ((4096.0 + (n_layers - 1) * 7680.0) if chained else (5120.0 + (n_layers - 1) * 10752.0)),
        # node-level contractions: the same 32 flop/B kernels on [N,128] / [3N,128] rows -> HBM-bound as well
        # per layer: W1,W2 fwd (2 x 1024 + silu' 512), Wu on 3N rows (3 x 1024), equiv bwd (3N rows x (512 x 4 + 512/3)), W2,W1 bwd
        # (1536 + 1536); head fwd 2560 + bwd 2560
        'node_gemm': N * (n_layers * (2560.0 + 3072.0 + 6656.0 + 3072.0) + 5120.0),
        'message': n_layers * (P * (512 + 8 + 80) + N * 512),
        'aggregate': (n_layers * (P * 2 * 512 + N * (512 * 2 + 1536) + 2 * P * 8 + P * 12) + (n_layers - 1) * (P * 512 + N * 1536 * 2)),
        'bwd_gather': (n_layers * (P * (1024 + 8 + 24) + N * 1536) + (n_layers - 1) * (P * 512 + N * 1536)),
        'bwd_message': n_layers * (P * (1024 + 8 + 160) + 2 * N * 512),
        'bwd_aggregate': (n_layers * (P * 512 + N * 1024 + 2 * P * 8) + (n_layers - 1) * (P * 512 + N * 1536 * 2)),
    }
    table = {}
    for name, total in per_step.items():
        if stage_n.get(name, 0) == 0 or stage_ms.get(name, 0.0) <= 0:
            continue
        per_launch = total * K / stage_n[name]
        avg_s = stage_ms[name] * 1e-3 / stage_n[name]
        ach = per_launch / avg_s * hbm['scale']
        table[name] = dict(bound='hbm', achieved=ach, peak=hbm['peak'], unit='GB/s', frac=ach / hbm['peak'],
                           launches=stage_n[name], ms_total=stage_ms[name], share=stage_ms[name] / dev_ms,
                           algorithmic_bytes_per_launch=per_launch)
    dominant = max(table, key=lambda k: table[k]['ms_total']) if table else None
    for name in ('nbr', 'geom', 'head', 'force', 'other'):      # latency-bound / small stages: time only
        if stage_n.get(name, 0) and stage_ms.get(name, 0.0) > 0:
            table[name] = dict(bound='latency', launches=stage_n[name], ms_total=stage_ms[name], share=stage_ms[name] / dev_ms)
    roofline = None
    if dominant:
        t = table[dominant]
        roofline = {'kernel': dominant + ('[tcgen05 3xTF32]' if backend in ('tc', 'ts') and 'gemm' in dominant else
                                          ('[fp32 SIMT]' if 'gemm' in dominant else '')),
                    'bound': 'hbm', 'achieved': t['achieved'], 'peak': t['peak'], 'unit': 'GB/s',
                    'frac': t['frac'], 'traffic': None, 'peak_source': pk['source'] + ' copy',
                    'share_of_step': t['share'], 'launches_timed': t['launches'],
                    'algorithmic_bytes_per_launch': t['algorithmic_bytes_per_launch']}
        if dominant == 'pair_gemm':   # the same launches as fp32-equivalent FLOP/s (one 2*128*128 product per row)
            n_prod = (4 + (n_layers - 1) * 8)
            roofline['tflops_fp32_equivalent'] = 2.0 * P * F * F * n_prod * K / (t['ms_total'] * 1e-3) * 1e-12
            roofline['tensor_pipe_passes'] = 3
            roofline['traffic'], roofline['traffic_note'] = traffic_record(workload)
    return table, roofline


def bench_single(ctx, workload, K, W, want_e2e=True, want_md=True):
    """One workload on this rank's GPU with the single-GPU path (every rank its own batch when world > 1)."""
    import torch
    from newtonnet_b200 import _lib as L
    from newtonnet_b200 import workloads
    from newtonnet_b200.engine import get_engine
    dev, lib, rank, world = ctx.dev, ctx.lib, ctx.rank, ctx.world
    z_h, pos_h, cell_h, batch_h = workloads.make(workload, seed=rank if workload in ('c1', 'c2') else 0)
    N, B = len(z_h), cell_h.shape[0]
    stress = workload in ('c3', 'c4')
    model, props = build_model(dev, stress)
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

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    lib.nn_profile_enable(1)
    lib.nn_launch_count(1)
    ctx.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(W, W + K):
        nl, out = step_device(i)
    ev1.record()
    torch.cuda.synchronize(); ctx.barrier()
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
    dev_ms = ctx.max_over_ranks(dev_ms)
    atoms_all = ctx.sum_over_ranks(float(N))
    value = atoms_all * K / (dev_ms * 1e-3)

    # ---- end to end through the public API with host buffers
    e2e = None
    if want_e2e:
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
        if workload == 'c3':
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
        ctx.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(W, W + K):
            step_e2e(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(e2e_s)
        h2d = z_h.nbytes + steps_pos[0].nbytes + cell_h.nbytes + batch_h.nbytes
        d2h = e_host.numel() * 4 + f_host.numel() * 4
        e2e = {'value': atoms_all * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_s / K * 1e3, 'api': api}

    # ---- device-resident MD (caller side): same box, integrator state in HBM, one CUDA graph per step
    md_info = None
    if want_md and workload in ('c1', 'c3') and world == 1:
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

    P = n_edges // 2
    stage_ms = {L.STAGES[k]: float(ms[k]) for k in range(12)}
    stage_n = {L.STAGES[k]: int(cnt[k]) for k in range(12)}
    table, roofline = roofline_from_stages(ctx, stage_ms, stage_n, dev_ms, K, N, P, pack.n_layers, workload)
    chained = ctx.backend == 'ts' and os.environ.get('NN_CHAIN', '1') != '0'
    res = {
        'value': value, 'ms_per_step': dev_ms / K, 'steps': K, 'warmup': W, 'scaling': 'weak',
        'config': {'workload': f'{workload}: {workloads.DESCRIPTION[workload]}', 'atoms_per_gpu': N,
                   'systems_per_gpu': B, 'directed_edges_per_gpu': n_edges, 'n_features': 128, 'n_basis': 20,
                   'n_interactions': pack.n_layers, 'cutoff': pack.cutoff, 'heads': props,
                   'gemm_backend': BACKENDS[ctx.backend] + (', chained two-CTA MLPs' if chained else ''),
                   'parallelism': (f'dp{world} (independent batches, no data-path collective)' if world > 1 else 'single GPU'),
                   'cache': 'per-step working set (pair tensors, %.1f GB) exceeds the 126 MB L2; positions change every step'
                            % (pack.n_layers * 5 * P * 512 / 1e9)},
        'gpu_launches': launches, 'clocks': clocks, 'e2e': e2e, 'roofline': roofline, 'stages': table,
    }
    if md_info:
        res['md_device_resident'] = md_info
    del pos_d, out, nl
    engine._nl = None; engine._ws = None
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- c4 across ranks: domain decomposition
def bench_c4_decomposed(ctx, K, W):
    """config 4: ONE 98,304-atom periodic box split into bricks - strong scaling: the total work is fixed."""
    import torch
    from newtonnet_b200 import _lib as L
    from newtonnet_b200 import workloads
    from newtonnet_b200.distributed import DomainDecomposition
    args, dev, lib, rank, world = ctx.args, ctx.dev, ctx.lib, ctx.rank, ctx.world
    z_h, pos_h, cell_h, batch_h = workloads.make('c4', seed=0)
    N = len(z_h)
    model, props = build_model(dev, stress=True)
    transport = os.environ.get('NN_DD_TRANSPORT', 'p2p')
    dd = DomainDecomposition(model, transport=transport, overlap=os.environ.get('NN_DD_OVERLAP', '1') != '0',
                             use_cuda_graph=os.environ.get('NN_DD_GRAPH', '1') != '0')
    rng = np.random.default_rng(100)
    steps_pos = [(pos_h + rng.normal(0, 0.01, pos_h.shape)).astype(np.float32) for _ in range(K + W)]
    z_d = torch.tensor(z_h, device=dev); cell_d = torch.tensor(cell_h, device=dev)
    pos_d = [torch.tensor(p, device=dev) for p in steps_pos]
    peer = transport == 'p2p'
    lib.nn_launch_count(1)
    out = dd(z_d, pos_d[0], cell_d)                        # plan + capacities + eager first step
    launches_per_step = int(lib.nn_launch_count(1))
    for i in range(W):
        out = dd(z_d, pos_d[i], cell_d, sync=not peer)
    if peer:
        dd.check()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    ctx.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(W, W + K):
        out = dd(z_d, pos_d[i], cell_d, sync=not peer)     # peer transport: no host synchronisation inside the timed region
    ev1.record()
    torch.cuda.synchronize(); ctx.barrier()
    dev_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    status = dd.check() if peer else None                  # stale plan / overflow / timeout in any timed step raises here
    clocks = sampler.stop() if rank == 0 else None
    f_last = out.gradient_force.float().cpu().numpy()
    e_last = float(out.energy.double().item())
    value = N * K / (dev_ms * 1e-3)

    # ---- end to end: pinned host positions -> DomainDecomposition -> energy / forces / stress on the host, every step
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory()
        pos_p = [pin(p) for p in steps_pos]
        e_host = torch.empty(1, dtype=torch.float32).pin_memory()
        f_host = torch.empty(N, 3, dtype=torch.float32).pin_memory()
        s_host = torch.empty(1, 3, 3, dtype=torch.float32).pin_memory()

        def step_e2e(i):
            pt = pos_p[i].to(dev, non_blocking=True)
            o = dd(z_d, pt, cell_d)                        # sync=True: status read every step
            e_host.copy_(o.energy, non_blocking=True); f_host.copy_(o.gradient_force, non_blocking=True)
            s_host.copy_(o.stress, non_blocking=True)
            torch.cuda.synchronize()
        for i in range(W):
            step_e2e(i)
        ctx.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(W, W + K):
            step_e2e(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(e2e_s)
        e2e = {'value': N * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(steps_pos[0].nbytes),
               'd2h_bytes_per_step': int(4 + N * 12 + 36), 'ms_per_step': e2e_s / K * 1e3,
               'api': 'newtonnet_b200.distributed.DomainDecomposition.__call__(z, pos, cell) on every rank: pinned host '
                      'positions in, energy / forces / stress copied to the host, status checked every step'}

    # ---- stage times of rank 0 from a few eager (non-graph) steps: CUDA events cannot be recorded inside a graph replay
    stage = None
    if peer:
        dd.use_cuda_graph = False
        lib.nn_profile_enable(1)
        torch.cuda.synchronize(); ctx.barrier()
        evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evp0.record()
        n_prof = 3
        for i in range(n_prof):
            dd(z_d, pos_d[W + i % K], cell_d, sync=False)
        evp1.record()
        torch.cuda.synchronize(); ctx.barrier()
        prof_ms = evp0.elapsed_time(evp1)
        ms = (C.c_float * 12)(); cnt = (C.c_int * 12)()
        lib.nn_profile_collect(ms, cnt, 12)
        lib.nn_profile_enable(0)
        dd.use_cuda_graph = True
        p = dd._peer
        n_local_edges = int(p.nl.status.cpu()[L.ST_N_EDGES])
        stage = (({L.STAGES[k]: float(ms[k]) for k in range(12)}, {L.STAGES[k]: int(cnt[k]) for k in range(12)}),
                 prof_ms, n_prof, p.n_owned, n_local_edges // 2)

    # ---- parity, computed in this run: N ranks vs 1 rank (CUDA) and vs the CPU oracle on a cluster
    parity = None
    if not args.no_parity:
        parity = {}
        if rank == 0:
            ref = model(z_d, pos_d[W + K - 1], cell_d, torch.zeros(N, dtype=torch.int64, device=dev))
            f1 = ref.gradient_force.float().cpu().numpy()
            e1 = float(ref.energy.double().item())
            parity['vs_1rank_cuda'] = {'max_abs_dF_eV_per_A': float(np.abs(f_last - f1).max()),
                                       'rel_dE': abs(e_last - e1) / abs(e1),
                                       'max_rel_dStress': float((out.stress - ref.stress).abs().max() / ref.stress.abs().max())}
            del ref
            from newtonnet_b200.engine import get_engine
            get_engine(dev)._ws = None; get_engine(dev)._nl = None; get_engine(dev)._graphs.clear()
            torch.cuda.empty_cache()
            sd = {k: v.detach().cpu().numpy() for k, v in seed0_weights().state_dict().items()}
            pos_last = steps_pos[W + K - 1]
            center = int(np.argmin(((pos_last - 0.5 * cell_h[0, 0, 0]) ** 2).sum(1)))
            t0 = time.perf_counter()
            from oracle.cluster import oracle_cluster_forces      # the checker, never the thing measured
            idx, f_or, info = oracle_cluster_forces(z_h, pos_last, cell_h, sd, center)
            info['seconds'] = time.perf_counter() - t0
            info['max_abs_dF_eV_per_A'] = float(np.abs(f_last[idx] - f_or).max())
            info['max_abs_F_eV_per_A'] = float(np.abs(f_or).max())
            info['what'] = ('oracle.radius_graph_cell_list + oracle.forward_analytic on a non-periodic cluster cut around the '
                            'destination atoms (receptive field of a force = 2 x 3 layers x 5 A), compared with the N-rank forces '
                            'of the same atoms')
            parity['vs_oracle_cluster'] = info
            parity['tolerance'] = {'forces_eV_per_A': 1e-4, 'energy_rel': 1e-5}
            parity['ok'] = bool(parity['vs_1rank_cuda']['max_abs_dF_eV_per_A'] <= 1e-4 and parity['vs_1rank_cuda']['rel_dE'] <= 1e-5
                                and info['max_abs_dF_eV_per_A'] <= 1e-4)
        ctx.barrier()

    p = dd._peer if peer else None
    halo_bytes, n_exchanges = p.halo_bytes_per_step() if p is not None else (None, None)
    halo_all = ctx.sum_over_ranks(float(halo_bytes)) if halo_bytes is not None else None
    res = {
        'value': value, 'ms_per_step': dev_ms / K, 'steps': K, 'warmup': W, 'scaling': 'strong',
        'config': {'workload': 'c4: ' + workloads.DESCRIPTION['c4'], 'atoms_total': N, 'heads': props,
                   'parallelism': f'spatial domain decomposition, {dd.plan.grid} bricks, ghost shell cutoff + skin = '
                                  f'{model.cutoff + dd.skin:.1f} A; halo exchange of ghost feature rows: {n_exchanges} exchanges per step, '
                                  + ('pack kernel storing into peer memory over NVLink (no NCCL on the data path), f_out / abar rows on a '
                                     'second stream behind node-level work; forces completed by owner-only peer stores; whole step = ONE '
                                     'CUDA graph, no host synchronisation in the timed region' if peer else
                                     'pack kernel + NCCL all_to_all, eager, all-reduce of forces/energy/virial'),
                   'owned_atoms_rank0': dd.plan.n_owned, 'ghost_atoms_rank0': dd.plan.n_ghost, 'plans_built': dd.n_plans,
                   'plan_skin_A': dd.skin, 'gemm_backend': BACKENDS[ctx.backend],
                   'cache': 'per-rank working set (pair tensors) exceeds the 126 MB L2; positions change every step',
                   'note': 'every step rebuilds the neighbour list on the device; the brick/ghost plan (host side) is reused while '
                           'no atom moved more than skin/2 (checked on the device every step); positions replicated on every rank, '
                           'results complete on every rank'},
        'gpu_launches': launches_per_step * K, 'gpu_launches_note': f'{launches_per_step} kernels per step (counted on the eager first '
                                                                    f'step) x {K} CUDA-graph replays',
        'clocks': clocks, 'e2e': e2e, 'parity': parity, 'roofline': None, 'stages': None,
    }
    if halo_bytes is not None:
        res['halo'] = {'bytes_per_step_rank0': int(halo_bytes), 'bytes_per_step_all_ranks': int(halo_all),
                       'exchanges_per_step': n_exchanges,
                       'nvlink_GBps_per_gpu_averaged_over_step': halo_bytes / (dev_ms / K * 1e-3) * 1e-9,
                       'status_words': status}
    if stage is not None and rank == 0:
        (stage_ms, stage_n), prof_ms, n_prof, n_owned, P_local = stage
        table, roofline = roofline_from_stages(ctx, stage_ms, stage_n, prof_ms, n_prof, n_owned, P_local, dd._peer.pack.n_layers, 'c4dd')
        if roofline:
            roofline['note'] = (f'rank 0, {n_prof} eager (non-graph) steps after the timed region with CUDA-event stage timers '
                                f'({prof_ms / n_prof:.2f} ms/step eager vs {dev_ms / K:.2f} ms/step graph replay)')
        res['roofline'], res['stages'] = roofline, table
    return res, dd


# ----------------------------------------------------------------------------- c5: training step
def bench_c5_training(ctx, K, W):
    """config 5: training step (energy + force loss, double backward, Adam) on MD17-shaped synthetic data,
    100 molecules x 21 atoms per GPU, data-parallel gradient all-reduce (one flat 1.6 MB bucket)."""
    import torch
    from newtonnet_b200 import workloads
    from newtonnet_b200.train import training_step
    dev, lib, rank, world = ctx.dev, ctx.lib, ctx.rank, ctx.world
    z, pos, cell, batch = workloads.make('c1', seed=rank)
    N = len(z)
    rng = np.random.default_rng(7 + rank)
    t = lambda a: torch.tensor(a, device=dev)
    e_t, f_t = t(rng.standard_normal(cell.shape[0]).astype(np.float32)), t(rng.standard_normal(pos.shape).astype(np.float32))
    model = seed0_weights().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=os.environ.get('NN_ADAM_FUSED', '1') == '1')
    args_t = (t(z), t(pos), t(cell), t(batch), e_t, f_t)
    graphed = os.environ.get('NN_TRAIN_GRAPH', '1') == '1'      # forward + double backward replayed as one CUDA graph
    if graphed:
        from newtonnet_b200.train import GraphedTrainingStep
        step = GraphedTrainingStep(model, opt, *args_t)
        training_step = lambda model, opt, *a: step(*a)     # noqa: E731,F811
    rng_p = np.random.default_rng(11 + rank)
    pos_steps = [t((pos + rng_p.normal(0, 0.01, pos.shape)).astype(np.float32)) for _ in range(K + W)]
    for i in range(W):
        loss = training_step(model, opt, args_t[0], pos_steps[i], *args_t[2:])
    lib.nn_launch_count(1)
    ctx.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        loss = training_step(model, opt, args_t[0], pos_steps[W + i], *args_t[2:])
    if graphed:
        step.settle()                                       # the last step's deferred overflow check, inside the timed region
    ev1.record()
    torch.cuda.synchronize(); ctx.barrier()
    dev_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    atoms_all = ctx.sum_over_ranks(float(N))
    launches = int(lib.nn_launch_count(1))
    if graphed:
        launches += step.kernels_per_replay * K                         # kernels of this library inside each graph replay
    v = atoms_all * K / (dev_ms * 1e-3)
    return {'metric': 'training ' + METRIC, 'value': v, 'unit': UNIT, 'ms_per_step': dev_ms / K, 'steps': K, 'warmup': W,
            'scaling': 'weak',
            'config': {'workload': 'c5: training step, 100 x 21 atoms per GPU, loss MSE(E) + 50 MSE(F), double backward, clip 1.0, '
                                   'Adam 1e-3 (torch fused kernel); new positions every step; '
                                   + (f'forward + backward replayed as one CUDA graph (GraphedTrainingStep, {step.nl.cap_edges} padded edge rows, weight gradients on a side branch, overflow flag checked one call later: {step.recaptures} re-captures)' if graphed
                                      else 'eager autograd (training_step)'), 'atoms_per_gpu': N,
                       'parallelism': f'dp{world}, one all-reduce of a flat 401,155-float gradient bucket',
                       'final_loss': float(loss)},
            'gpu_launches': launches}


# ----------------------------------------------------------------------------- this repo's arm
def main_b200(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1 and args.gpus > 1:       # convenience: relaunch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29511', os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd, stdout=_REAL_STDOUT)     # the children write the JSON line to the real stdout
    import torch
    import torch.distributed as dist
    ctx = Ctx(args)
    rank = ctx.rank
    K, W = ctx.K, ctx.W
    wl = args.workload
    dd = None
    if wl == 'c5':
        head = bench_c5_training(ctx, K, W)
        head.update({'e2e': {'value': head['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                             'note': 'training data resident on the device'}, 'roofline': None, 'clocks': None})
    elif wl == 'c4' and world > 1:
        head, dd = bench_c4_decomposed(ctx, K, W)
    else:
        head = bench_single(ctx, wl, K, W, want_e2e=not args.no_e2e)

    extra = {}
    if not args.no_extra and wl == 'c4':
        def guarded(name, fn):
            try:
                extra[name] = fn()
            except Exception as exc:      # noqa: BLE001 - an extra block must not take the headline down
                extra[name] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
            torch.cuda.empty_cache()
        Kx = min(K, 10)
        guarded('c2', lambda: bench_single(ctx, 'c2', Kx, W, want_e2e=(world == 1 and not args.no_e2e), want_md=False))
        if world == 1:
            guarded('c3', lambda: bench_single(ctx, 'c3', max(K, 20), W, want_e2e=not args.no_e2e))
        guarded('c5', lambda: bench_c5_training(ctx, Kx, W))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _, _ = run_oracle_timed(wl, steps=40, warmup=1, budget_s=15.0)   # ~15 s of CPU work
    if dd is not None:
        dd.close()
    if rank == 0:
        line = {
            'metric': head.get('metric', METRIC), 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': head['steps'],
            'warmup': head['warmup'], 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': head['scaling'],
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': head['config'],
            'gpu_launches': head['gpu_launches'], 'clocks': head.get('clocks'), 'e2e': head.get('e2e'),
            'roofline': head.get('roofline'), 'cpu_baseline': cpu_base, 'stages': head.get('stages'),
        }
        for k in ('parity', 'halo', 'gpu_launches_note', 'md_device_resident'):
            if head.get(k) is not None:
                line[k] = head[k]
        if extra:
            line['extra'] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None

if __name__ == '__main__':
    # stdout carries exactly one JSON line: anything a library prints there (e.g. NCCL's version banner)
    # is rerouted to stderr, the result line goes to the saved descriptor
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = _REAL_STDOUT
    a = parse()
    sys.exit(main_reference(a) if a.impl == 'reference' else main_b200(a))
