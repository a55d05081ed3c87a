"""String -> dtype; mirrors newtonnet/layers/precision.py:3-13.  The kernels compute in fp32; fp64/fp16
models are accepted and cast at the boundary."""
import torch

__all__ = ['get_precision_by_string']


def get_precision_by_string(key):
    table = {'float32': torch.float32, 'float': torch.float32, 'single': torch.float32,
             'float64': torch.float64, 'double': torch.float64,
             'float16': torch.float16, 'half': torch.float16}
    if key not in table:
        raise ValueError(f'precision {key} is not supported')
    return table[key]
