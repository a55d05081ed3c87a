"""The example drivers (examples/) run end to end on the GPU: reference-format checkpoint + extxyz in, trajectory / training out."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_weights

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'examples'))


def _files(tmp_path, n_frames):
    from newtonnet_b200 import data as D
    from newtonnet_b200.compat import model_from_state_dict
    kat = np.load(f'{GOLDEN}/md17_kat.npz')
    ckpt = tmp_path / 'best_model.pt'
    torch.save(model_from_state_dict({k: torch.tensor(v).double() for k, v in load_weights('md17').items()}), ckpt)
    xyz = tmp_path / 'frames.xyz'
    D.write_extxyz(str(xyz), [{'z': kat['numbers'], 'pos': kat['positions'][k], 'energy': kat['energy'][k], 'force': kat['forces'][k]}
                              for k in range(n_frames)])
    return str(ckpt), str(xyz), kat


def test_md_example(tmp_path):
    import md_aspirin
    from newtonnet_b200 import data as D
    ckpt, xyz, kat = _files(tmp_path, 2)
    out = str(tmp_path / 'traj.xyz')
    md = md_aspirin.main(ckpt, xyz, steps=40, out=out, log_interval=10)
    assert md.step == 40
    traj = D.read_extxyz(out)
    assert len(traj) == 4 and traj[0]['z'].tolist() == kat['numbers'].tolist()
    np.testing.assert_allclose(traj[-1]['pos'], md.positions, atol=1e-8)
    assert 50.0 < md.temperature()[0] < 1500.0          # 21 atoms: the instantaneous temperature fluctuates a lot


def test_training_example(tmp_path, monkeypatch):
    import train_md17
    ckpt, xyz, kat = _files(tmp_path, 24)
    monkeypatch.chdir(tmp_path)
    train_md17.main(xyz, epochs=2, batch_size=8)
    from newtonnet_b200.compat import load_model
    model = load_model(str(tmp_path / 'train_state.pt'), map_location='cuda:0')       # the layout the reference's trainer writes
    assert model.output_properties == ['energy', 'gradient_force']
