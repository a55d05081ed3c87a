"""Operator-level profile of one eager c5 training step (torch.profiler): which ATen ops / custom Functions the ~900 kernels
of the step come from.  Diagnostic only - timings under the profiler are not bench values.
    python tools/train_profile.py > gpurun_out/train_profile.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from newtonnet_b200 import workloads  # noqa: E402
from newtonnet_b200.train import training_step  # noqa: E402

dev = torch.device('cuda:0')
z, pos, cell, batch = workloads.make('c1', seed=0)
rng = np.random.default_rng(7)
t = lambda a: torch.tensor(a, device=dev)
e_t, f_t = t(rng.standard_normal(cell.shape[0]).astype(np.float32)), t(rng.standard_normal(pos.shape).astype(np.float32))
model = bench.seed0_weights().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
args = (t(z), t(pos), t(cell), t(batch), e_t, f_t)
for _ in range(3):
    training_step(model, opt, *args)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            record_shapes=True) as prof:
    training_step(model, opt, *args)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by='self_cuda_time_total', row_limit=70, max_name_column_width=48,
                                                         max_shapes_column_width=70))
