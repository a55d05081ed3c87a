"""Kernel timeline of the domain-decomposed c4 step on rank 0 (torch.profiler / CUPTI; nsys is not in the image).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dd_profile.py [out.json]
Prints per-kernel totals over `STEPS` CUDA-graph replays and the GPU idle share, writes them as JSON."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from newtonnet_b200 import workloads  # noqa: E402
from newtonnet_b200.distributed import DomainDecomposition  # noqa: E402

STEPS = 5
rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
z, pos, cell, batch = workloads.make('c4', 0)
model, _ = bench.build_model(dev, True)
dd = DomainDecomposition(model, overlap=os.environ.get('NN_DD_OVERLAP', '1') != '0')
t = lambda a: torch.tensor(a, device=dev)
rng = np.random.default_rng(1)
ps = [t((pos + rng.normal(0, 0.01, pos.shape)).astype(np.float32)) for _ in range(STEPS + 4)]
zt, ct = t(z), t(cell)
for i in range(4):
    dd(zt, ps[i], ct, sync=False)
dd.check()
torch.cuda.synchronize(); dist.barrier(device_ids=[local])
from torch.profiler import ProfilerActivity, profile
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ev0.record()
    for i in range(STEPS):
        dd(zt, ps[4 + i], ct, sync=False)
    ev1.record()
    torch.cuda.synchronize()
dist.barrier(device_ids=[local])
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = {}
    for e in evs:
        k = e.name.split('(')[0][:60]
        a = tot.setdefault(k, [0.0, 0])
        a[0] += e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total
        a[1] += 1
    wall = ev0.elapsed_time(ev1) * 1e3
    busy = sum(v[0] for v in tot.values())
    rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
    print(f'rank 0: {wall / STEPS:.1f} us/step wall, sum of kernel times {busy / STEPS:.1f} us/step (both streams)')
    for k, (us, n) in rows[:40]:
        print(f'{us / STEPS:9.1f} us/step  {n / STEPS:6.1f} launches/step  {us / n:8.1f} us avg  {k}')
    # serialized timeline gaps on the union of intervals
    iv = sorted((e.time_range.start, e.time_range.end) for e in evs)
    union, cur_s, cur_e = 0.0, None, None
    for s, e_ in iv:
        if cur_s is None:
            cur_s, cur_e = s, e_
        elif s <= cur_e:
            cur_e = max(cur_e, e_)
        else:
            union += cur_e - cur_s; cur_s, cur_e = s, e_
    union += cur_e - cur_s
    print(f'GPU busy (union of kernel intervals) {union / STEPS:.1f} us/step -> idle {100 * (1 - union / wall):.1f} %')
    out = sys.argv[1] if len(sys.argv) > 1 else None
    if out:
        json.dump({'us_per_step_wall': wall / STEPS, 'us_per_step_busy_union': union / STEPS, 'world': dist.get_world_size(),
                   'kernels': [{'name': k, 'us_per_step': us / STEPS, 'launches_per_step': n / STEPS} for k, (us, n) in rows]},
                  open(out, 'w'), indent=1)
dd.close()
dist.destroy_process_group()
