"""Output heads, scalers' companions and the result bag; mirrors newtonnet/models/output.py.

Supported heads (the hot path of BASELINE.json): 'energy', 'gradient_force', 'stress', 'virial', plus
'direct_force' and 'hessian' (SURVEY.md section 8f rank 2).  'charge', 'bec' raise NotImplementedError
(out of scope; charge/bec need the un-vendored `les` package).

The heads own parameters and flags only.  Energies, forces and virials are produced together by one
nn_eval call (csrc/eval.cu); NewtonNet.forward distributes the results to the bag in the order of
`output_properties`, as models/newtonnet.py:98-102 does.
"""
import torch
from torch import nn

__all__ = ['get_output_by_string', 'get_aggregator_by_string', 'CustomOutputSet', 'DirectProperty',
           'DerivativeProperty', 'SecondDerivativeProperty', 'EnergyOutput', 'GradientForceOutput', 'DirectForceOutput',
           'VirialOutput', 'StressOutput', 'HessianOutput', 'EnergyAggregator', 'NullAggregator', 'SumAggregator']

_UNSUPPORTED = ('charge', 'bec')


def get_output_by_string(key, n_features=None, activation=None):
    if key == 'energy':
        return EnergyOutput(n_features, activation)
    if key == 'gradient_force':
        return GradientForceOutput()
    if key == 'direct_force':
        return DirectForceOutput(n_features, activation)
    if key == 'virial':
        return VirialOutput()
    if key == 'stress':
        return StressOutput()
    if key == 'hessian':
        return HessianOutput()
    if key in _UNSUPPORTED:
        raise NotImplementedError(f"output '{key}' is outside the B200 energy/force/stress path")
    raise NotImplementedError(f'Output type {key} is not implemented yet')


def get_aggregator_by_string(key):
    if key == 'energy':
        return EnergyAggregator()
    if key in ('gradient_force', 'direct_force', 'virial', 'stress', 'hessian'):
        return NullAggregator()
    if key in _UNSUPPORTED:
        raise NotImplementedError(f"output '{key}' is outside the B200 energy/force/stress path")
    raise NotImplementedError(f'Aggregate type {key} is not implemented yet')


class CustomOutputSet:
    """Attribute bag (models/output.py:51-54).  Values given as zero-argument callables under `_lazy`
    are materialised on first access (edge_index needs the edge count on the host)."""

    def __init__(self, _lazy=None, **outputs):
        object.__setattr__(self, '_lazy', dict(_lazy or {}))
        for key, value in outputs.items():
            setattr(self, key, value)

    def __getattr__(self, name):
        lazy = object.__getattribute__(self, '_lazy')
        if name in lazy:
            value = lazy.pop(name)()
            setattr(self, name, value)
            return value
        raise AttributeError(name)


class DirectProperty(nn.Module):
    pass


class DerivativeProperty(nn.Module):
    def __init__(self):
        super().__init__()
        self.create_graph = False   # set by NewtonNet.train() / the calculator (models/output.py:64)


class SecondDerivativeProperty(DerivativeProperty):
    pass


class EnergyOutput(DirectProperty):
    """128 -> 128 -> 128 -> 1 MLP with SiLU (models/output.py:90-96); parameters only."""

    def __init__(self, n_features, activation):
        super().__init__()
        act = activation if activation is not None else nn.SiLU()
        self.layers = nn.Sequential(
            nn.Linear(n_features, n_features), act,
            nn.Linear(n_features, n_features), act,
            nn.Linear(n_features, 1),
        )


class DirectForceOutput(DirectProperty):
    """force_i[c] = sum_f MLP(atom_node_i)[f] * force_node_i[c][f]  (models/output.py:115-132); parameters only,
    evaluated by nn_eval (three GEMMs + k_direct_force)."""

    def __init__(self, n_features, activation):
        super().__init__()
        if n_features is None:
            raise ValueError("get_output_by_string('direct_force') needs n_features")
        act = activation if activation is not None else nn.SiLU()
        self.layers = nn.Sequential(
            nn.Linear(n_features, n_features), act,
            nn.Linear(n_features, n_features), act,
            nn.Linear(n_features, n_features),
        )


class GradientForceOutput(DerivativeProperty):
    """force = -dE/dpos (models/output.py:109-113)."""


class HessianOutput(SecondDerivativeProperty):
    """hessian[i,a,j,b] = d^2 E / d pos_ia d pos_jb (models/output.py:134-152); evaluated on the differentiable
    path (newtonnet_b200/train.py): one reverse pass through the force graph per row."""


class VirialOutput(DerivativeProperty):
    """virial = -dE/dD (models/output.py:161-165)."""


class StressOutput(DerivativeProperty):
    """stress = dE/dD / det(cell) (models/output.py:174-180)."""


class EnergyAggregator(nn.Module):
    """Per-system sum of atomic energies (models/output.py:246-247); done in fp64, fixed order, inside
    csrc/pair_ops.cu:k_energy_sum.  The latent-Ewald branch (:234-244) needs `les` and is out of scope."""


class NullAggregator(nn.Module):
    def forward(self, output, outputs):
        return output


class SumAggregator(nn.Module):
    """Legacy class name found in the shipped checkpoint scripts/md17_model/.../best_model.pt."""
