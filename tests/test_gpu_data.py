"""Data side feeding the training step on the GPU: extxyz file -> statistics -> scalers -> training_step."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_xyz_to_training_steps(tmp_path):
    from newtonnet_b200 import data as D
    from newtonnet_b200.models import NewtonNet
    from newtonnet_b200.train import training_step
    kat = np.load(f'{GOLDEN}/md17_kat.npz')
    path = tmp_path / 'train.xyz'
    with open(path, 'w') as fh:                      # the layout of scripts/md17_data/*/raw/*.xyz
        for k in range(0, 64):
            fh.write('21\nProperties=species:S:1:pos:R:3:forces:R:3 energy=%.8f pbc="F F F"\n' % kat['energy'][k])
            for zi, p, f in zip(kat['numbers'], kat['positions'][k], kat['forces'][k]):
                fh.write('%s %.8f %.8f %.8f %.8f %.8f %.8f\n' % (D.SYMBOLS[zi], *p, *f))
    frames = D.read_extxyz(str(path))
    assert len(frames) == 64 and frames[0]['z'].tolist() == kat['numbers'].tolist()
    stats = D.molecular_statistics(frames)
    torch.manual_seed(0)
    model = NewtonNet(output_properties=['energy', 'gradient_force']).to('cuda:0')
    D.fit_scalers(model, stats)
    # one composition only: the minimum-norm least-squares shifts reproduce the mean energy exactly
    e_mean = np.mean([f['energy'] for f in frames])
    shift = model.scalers[0].shift.weight.detach().cpu().numpy()[:, 0]
    assert abs(shift[kat['numbers']].sum() - e_mean) < 1e-2
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batch = D.collate(frames[:16], device='cuda:0')
    losses = [training_step(model, opt, *batch).item() for _ in range(6)]
    assert all(np.isfinite(losses)) and min(losses[1:]) < losses[0]
    # with the fitted shift the initial energy error is of the order of the residual scale, not of |E| ~ 1.8e4 eV
    assert losses[0] < 1e3
