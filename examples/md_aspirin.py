"""Aspirin MD as in the reference's scripts/simulate.py (Langevin, 0.5 fs, 300 K, friction 1/(500 fs)) - with the
integrator state resident on the GPU: one CUDA-graph replay per time step, no per-step host round trip.

    python examples/md_aspirin.py CHECKPOINT.pt FRAMES.xyz [steps] [out.xyz]

CHECKPOINT.pt: a reference checkpoint (pickled module or state dict), e.g. md17_model/training_1/models/best_model.pt;
FRAMES.xyz: extended xyz, the first frame is the start (e.g. md17_data/aspirin/ccsd_test/raw/aspirin_ccsd-test.xyz).
"""
import sys
import time

import numpy as np

from newtonnet_b200 import data
from newtonnet_b200.md import FS, DeviceMD
from newtonnet_b200.utils.ase_interface import MLAseCalculator


def main(checkpoint, xyz, steps=20000, out=None, log_interval=100):
    frame = data.read_extxyz(xyz, limit=1)[0]
    calc = MLAseCalculator(model_path=checkpoint, properties=['energy', 'forces'], precision='single', device='cuda')
    md = DeviceMD(calc.model, frame['z'], frame['pos'], cell=frame['cell'][None], timestep=0.5 * FS, temperature_K=300.0,
                  friction=1.0 / (500 * FS), seed=0)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        n = min(log_interval * 10, steps - done)
        log = md.run(n, trajectory_interval=log_interval)
        done += n
        e, k = log['energy'][-1, 0], log['kinetic'][-1, 0]
        print(f'step {done:8d}  Epot {e:14.6f} eV  Ekin {k:10.6f} eV  T {md.temperature()[0]:7.1f} K  '
              f'{(time.perf_counter() - t0) / done * 1e3:.3f} ms/step', flush=True)
        if out:
            data.write_extxyz(out, [{'z': frame['z'], 'pos': p, 'cell': frame['cell']} for p in log['positions']], append=True)
    return md


if __name__ == '__main__':
    a = sys.argv[1:]
    if len(a) < 2:
        sys.exit(__doc__)
    main(a[0], a[1], int(a[2]) if len(a) > 2 else 20000, a[3] if len(a) > 3 else None)
