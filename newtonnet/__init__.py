"""Drop-in import path: `newtonnet` re-exports `newtonnet_b200` (the B200 CUDA implementation of NewtonNet's energy /
force / stress path) under the reference's module names, so reference-side code keeps its import lines
(scripts/simulate.py:6 `from newtonnet.utils.ase_interface import MLAseCalculator`, scripts/newtonnet_train.py:9
`from newtonnet.models import NewtonNet`, the calculator's own imports utils/ase_interface.py:8-13) and the
reference's whole-module pickles (`torch.load(..., weights_only=False)`, utils/ase_interface.py:87) resolve their
`newtonnet.models.newtonnet.NewtonNet` ... class paths to this package's classes.

Only the hot path is mirrored: newtonnet.train (Trainer, wandb loop) and the PyG dataset classes
of newtonnet.data are outside it; importing them raises with a pointer to what exists instead.
"""
__name__ = 'NewtonNet'
__version__ = '2.1.0'
__backend__ = 'newtonnet_b200'
