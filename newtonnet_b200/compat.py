"""Loading reference checkpoints.

The reference saves whole pickled modules (newtonnet/train/trainer.py:219,221) and the calculator loads
them with torch.load(weights_only=False) (newtonnet/utils/ase_interface.py:87).  Those pickles name
classes under `newtonnet.*`; `load_model` resolves them to this package's mirror classes (or to the real
reference classes when that package is importable), then rebuilds a fresh `newtonnet_b200.NewtonNet`
from the state dict, so old layouts load too: the shipped MD17 checkpoint
(scripts/md17_model/training_1/models/best_model.pt) uses `embedding_layer` (singular),
`infer_properties` and `SumAggregator`.
"""
import contextlib
import importlib
import pickle
import sys
import types

import torch

_ALIASES = {
    'newtonnet': 'newtonnet_b200',
    'newtonnet.models': 'newtonnet_b200.models',
    'newtonnet.models.newtonnet': 'newtonnet_b200.models.newtonnet',
    'newtonnet.models.output': 'newtonnet_b200.models.output',
    'newtonnet.layers': 'newtonnet_b200.layers',
    'newtonnet.layers.representations': 'newtonnet_b200.layers.representations',
    'newtonnet.layers.scalers': 'newtonnet_b200.layers.scalers',
    'newtonnet.layers.activations': 'newtonnet_b200.layers.activations',
    'newtonnet.layers.precision': 'newtonnet_b200.layers.precision',
}

_LEGACY_KEYS = {
    'embedding_layer.node_embedding.weight': 'embedding_layers.node_embedding.weight',
    'embedding_layer.edge_embedding.frequencies': 'embedding_layers.edge_embedding.embedding.frequencies',
}


@contextlib.contextmanager
def reference_class_aliases():
    """Make `newtonnet.*` importable as this package while unpickling (no-op if a `newtonnet` package - the real one
    or this repo's shim - is already loaded or importable)."""
    added = []
    if 'newtonnet' not in sys.modules:
        try:
            importlib.import_module('newtonnet.models.newtonnet')
        except ImportError:
            for ref in [m for m in sys.modules if m == 'newtonnet' or m.startswith('newtonnet.')]:
                sys.modules.pop(ref, None)           # a half-imported package must not shadow the aliases
            for ref, ours in _ALIASES.items():
                sys.modules[ref] = importlib.import_module(ours)
                added.append(ref)
    try:
        yield
    finally:
        for ref in added:
            sys.modules.pop(ref, None)


class _TolerantUnpickler(pickle.Unpickler):
    """Whole-module pickles of the current reference contain a `les.Les` instance (and its submodules) inside
    aggregators.N (models/output.py:229); `les` is an un-vendored dependency.  Classes of packages that are not on the
    hot path and cannot be imported resolve to inert nn.Module stand-ins - `convert_state_dict` drops aggregators.*."""
    _SKIPPABLE = ('les', 'torch_geometric', 'wandb')

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            if module.split('.')[0] in self._SKIPPABLE:
                return type(name, (torch.nn.Module,), {'__module__': module, '_newtonnet_b200_stub': True})
            raise


def _tolerant_pickle_module():
    m = types.ModuleType('newtonnet_b200_tolerant_pickle')
    m.__dict__.update({k: v for k, v in pickle.__dict__.items() if not k.startswith('__')})
    m.Unpickler = _TolerantUnpickler
    m.load = lambda f, **kw: _TolerantUnpickler(f, **kw).load()
    return m


def convert_state_dict(sd):
    """Rename legacy keys; drop entries that are not parameters of the supported path."""
    out = {}
    for k, v in sd.items():
        k = _LEGACY_KEYS.get(k, k)
        if k.startswith('aggregators.'):
            continue
        out[k] = v
    return out


def _find_cutoff(obj, default=5.0):
    for path in ('embedding_layers.edge_embedding.norm', 'embedding_layers.edge_embedding.radius_graph',
                 'embedding_layer.norm'):
        cur = obj
        try:
            for name in path.split('.'):
                cur = getattr(cur, name)
            return float(cur.r)
        except AttributeError:
            continue
    return default


def _activation_name(obj):
    """Name of the activation of a pickled reference module (layers/activations.py:5-31), read off the module tree."""
    act = None
    try:
        act = obj.interaction_layers[0].message_nodepart[1]
    except (AttributeError, IndexError, TypeError):
        return None
    name = type(act).__name__
    return {'SiLU': 'swish', 'ReLU': 'relu', 'GELU': 'gelu', 'Tanh': 'tanh', 'Softplus': 'ssp', 'ShiftedSoftplus': 'ssp'}.get(name, name)


def model_from_state_dict(sd, output_properties=None, cutoff=5.0, activation='swish'):
    """activation: the reference's activation key; a state dict does not record it, so callers that know it pass it.
    Anything but swish / silu raises (the kernels implement SiLU) instead of silently evaluating the wrong function."""
    from newtonnet_b200.models.newtonnet import NewtonNet
    sd = convert_state_dict(sd)
    emb = sd['embedding_layers.node_embedding.weight']
    n_features = emb.shape[1]
    n_basis = sd['embedding_layers.edge_embedding.embedding.frequencies'].numel()
    n_int = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('interaction_layers.'))
    layer_norm = any('.layer_norm.' in k for k in sd)
    if output_properties is None:
        output_properties = ['energy', 'gradient_force']
    model = NewtonNet(cutoff=cutoff, n_features=n_features, n_basis=n_basis, n_interactions=n_int,
                      activation=activation, layer_norm=layer_norm, output_properties=list(output_properties))
    model = model.to(emb.dtype)
    model.load_state_dict(sd, strict=True)
    return model


def load_model(path_or_obj, map_location=None):
    """Checkpoint path / pickled module / state dict / module  ->  newtonnet_b200.NewtonNet."""
    from newtonnet_b200.models.newtonnet import NewtonNet
    obj = path_or_obj
    if isinstance(obj, (str, bytes)) or hasattr(obj, 'read') or hasattr(obj, '__fspath__'):
        with reference_class_aliases():
            obj = torch.load(obj, map_location=map_location, weights_only=False, pickle_module=_tolerant_pickle_module())
    if isinstance(obj, dict) and 'model_state_dict' in obj:
        obj = obj['model_state_dict']       # train_state.pt layout (trainer.py:242-251)
    if isinstance(obj, dict):
        model = model_from_state_dict(obj)
    elif isinstance(obj, NewtonNet) and hasattr(obj, 'embedding_layers') and hasattr(obj, 'return_node_features'):
        model = obj
    elif isinstance(obj, torch.nn.Module):
        props = getattr(obj, 'output_properties', None)
        if props is None:
            props = getattr(obj, 'infer_properties', None)
        # bypass NewtonNet.state_dict() of half-initialised mirror objects: walk the raw module tree
        sd = torch.nn.Module.state_dict(obj)
        act = _activation_name(obj)
        model = model_from_state_dict(sd, output_properties=props, cutoff=_find_cutoff(obj), activation=act or 'swish')
    else:
        raise TypeError(f'cannot build a NewtonNet from {type(obj)}')
    if map_location is not None:
        model = model.to(map_location)
    return model
