"""Minimal MD17 training loop on the B200 path - what the reference's scripts/newtonnet_train.py + train/trainer.py do
per step (statistics -> scalers, energy + force loss, double backward, clip, Adam), without ASE / PyG / wandb.
Data-parallel under torchrun (one process per GPU; gradients averaged with ONE all-reduce of a flat bucket).

    python examples/train_md17.py TRAIN.xyz [epochs] [batch_size]
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 examples/train_md17.py TRAIN.xyz
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from newtonnet_b200 import data
from newtonnet_b200.models import NewtonNet
from newtonnet_b200.train import training_step


def main(xyz, epochs=10, batch_size=32):
    world, rank = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    frames = data.read_extxyz(xyz)
    torch.manual_seed(0)                                           # same initial weights on every rank
    model = NewtonNet(output_properties=['energy', 'gradient_force']).to(dev)
    data.fit_scalers(model, data.molecular_statistics(frames))     # scripts/newtonnet_train.py:88-90
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    rng = np.random.default_rng(0)
    for epoch in range(epochs):
        order = rng.permutation(len(frames))
        losses = []
        for b0 in range(0, len(order) - world * batch_size + 1, world * batch_size):
            mine = order[b0 + rank * batch_size: b0 + (rank + 1) * batch_size]
            batch = data.collate([frames[i] for i in mine], device=dev)
            losses.append(training_step(model, opt, *batch, force_weight=50.0, clip_grad=1.0))
        if rank == 0 and losses:
            print(f'epoch {epoch:4d}  loss {torch.stack(losses).mean().item():.6f}', flush=True)
    if rank == 0:
        torch.save({'model_state_dict': model.state_dict()}, 'train_state.pt')     # layout of train/trainer.py:242-251
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = sys.argv[1:]
    if not a:
        sys.exit(__doc__)
    main(a[0], int(a[1]) if len(a) > 1 else 10, int(a[2]) if len(a) > 2 else 32)
