"""Per-element scale / shift; mirrors newtonnet/layers/scalers.py (parameter names and factory keys).

The arithmetic (e_i * scale[z_i] + shift[z_i], scalers.py:55-58) runs inside the fused energy-head
kernel (csrc/pair_ops.cu: k_energy_atom); this module only owns the parameters.
"""
import torch
from torch import nn

__all__ = ['get_scaler_by_string', 'set_scaler_by_string', 'ScaleShift']

_SCALER_SPEC = {   # key -> (scale, shift) initial values, None = absent (scalers.py:5-24)
    'energy': (1.0, 0.0), 'gradient_force': (None, None), 'direct_force': (1.0, None), 'hessian': (None, None),
    'virial': (None, None), 'stress': (None, None), 'charge': (0.1, 0.0), 'bec': (None, None),
}


def get_scaler_by_string(key):
    if key not in _SCALER_SPEC:
        raise NotImplementedError(f'Scaler type {key} is not implemented yet')
    scale, shift = _SCALER_SPEC[key]
    return ScaleShift(scale=scale, shift=shift)


def set_scaler_by_string(key, scaler, stats, fit_scale=True, fit_shift=True):
    if scaler.scale is not None and key in stats and fit_scale:
        scaler.set_scale(stats[key]['scale'])
    if scaler.shift is not None and key in stats and fit_shift:
        scaler.set_shift(stats[key]['shift'])
    return scaler


class ScaleShift(nn.Module):
    """Node-level scale and shift (Embedding(119, 1) each, row 0 = padding)."""

    def __init__(self, scale=None, shift=None):
        super().__init__()
        # the reference initialises with ones / zeros whatever value is passed (scalers.py:44-45)
        self.scale = None if scale is None else nn.Embedding.from_pretrained(torch.ones(119, 1), freeze=False, padding_idx=0)
        self.shift = None if shift is None else nn.Embedding.from_pretrained(torch.zeros(119, 1), freeze=False, padding_idx=0)

    def forward(self, output, outputs):
        if self.scale is not None:
            output = output * self.scale(outputs.z)
        if self.shift is not None:
            output = output + self.shift(outputs.z)
        return output

    def set_scale(self, scale):
        self.scale.weight.data = scale.reshape(-1, 1)

    def set_shift(self, shift):
        self.shift.weight.data = shift.reshape(-1, 1)

    def __repr__(self):
        return f'{self.__class__.__name__}(scale={self.scale is not None}, shift={self.shift is not None})'
