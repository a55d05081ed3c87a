"""String -> activation factory; mirrors newtonnet/layers/activations.py:5-31.

The CUDA kernels implement SiLU ('swish' / 'silu', the reference default, scripts/config.yml:34); every
other key of the reference raises here instead of silently running something else.
"""
from torch import nn

__all__ = ['get_activation_by_string']


def get_activation_by_string(key):
    if key in ('swish', 'silu'):
        return nn.SiLU()
    if key in ('relu', 'elu', 'leaky_relu', 'tanh', 'sigmoid', 'softplus', 'gelu', 'ssp', 'swiglu'):
        raise NotImplementedError(f"activation '{key}' is not implemented by the B200 kernels (SiLU only)")
    raise NotImplementedError("The activation function '%s' is unknown." % str(key))
