"""NewtonNet model - host-side mirror of newtonnet/models/newtonnet.py.

Same constructor, module tree and state-dict keys as the reference, so reference
checkpoints load unchanged; `forward(z, pos, cell, batch)` returns the same attribute bag but is computed
by the CUDA library (one nn_nbr_count/nn_nbr_fill + one nn_eval), not by PyTorch ops.
"""
import torch
from torch import nn

from newtonnet_b200.layers.activations import get_activation_by_string
from newtonnet_b200.layers.representations import EdgeEmbedding
from newtonnet_b200.layers.scalers import get_scaler_by_string
from newtonnet_b200.models.output import (CustomOutputSet, DerivativeProperty, get_aggregator_by_string,
                                          get_output_by_string)

__all__ = ['NewtonNet', 'EmbeddingNet', 'InteractionNet']

_SUPPORTED = ('energy', 'gradient_force', 'stress', 'virial', 'direct_force', 'hessian')


class NewtonNet(nn.Module):
    """Molecular Newtonian message passing (reference models/newtonnet.py:12-71).

    Parameters are those of the reference: cutoff, n_features, n_basis, n_interactions, activation,
    layer_norm, output_properties.  The kernels are specialised for n_features=128, n_basis=20 and SiLU
    (scripts/config.yml:30-35); other values raise at the first forward.
    """

    def __init__(self, cutoff: float = 5.0, n_features: int = 128, n_basis: int = 20, n_interactions: int = 3,
                 activation: str = 'swish', layer_norm: bool = False, output_properties: list = []) -> None:
        super().__init__()
        activation = get_activation_by_string(activation)
        self.embedding_layers = EmbeddingNet(cutoff=cutoff, n_features=n_features, n_basis=n_basis)
        self.interaction_layers = nn.ModuleList([
            InteractionNet(n_features=n_features, n_basis=n_basis, activation=activation, layer_norm=layer_norm)
            for _ in range(n_interactions)])
        self.output_properties = output_properties
        self.output_layers = nn.ModuleList()
        self.scalers = nn.ModuleList()
        self.aggregators = nn.ModuleList()
        for key in self.output_properties:
            output_layer = get_output_by_string(key, n_features, activation)
            self.output_layers.append(output_layer)
            if isinstance(output_layer, DerivativeProperty):
                self.embedding_layers.requires_dr = True
            self.scalers.append(get_scaler_by_string(key))
            self.aggregators.append(get_aggregator_by_string(key))
        self.return_node_features = True
        self._pack = None
        self._pack_key = None

    # ------------------------------------------------------------------ weights
    @property
    def cutoff(self):
        return float(self.embedding_layers.edge_embedding.radius_graph.r)

    def _weight_pack(self, device):
        from newtonnet_b200.engine import WeightPack
        params = list(self.state_dict(keep_vars=True).items())
        key = (str(device), self.cutoff) + tuple((k, t.data_ptr(), t._version) for k, t in params)
        if self._pack is None or self._pack_key != key:
            sd = {k: t for k, t in params}
            head = [i for i, k in enumerate(self.output_properties) if k == 'energy']
            if not head:
                raise RuntimeError("output_properties must contain 'energy' (models/newtonnet.py:98-102 evaluates "
                                   "the heads in order and every other supported head differentiates the energy)")
            i = head[0]
            # the pack reads head / scaler parameters under index 0
            remap = {}
            d = [j for j, k in enumerate(self.output_properties) if k == 'direct_force']
            for k, t in sd.items():
                if k.startswith(f'output_layers.{i}.'):
                    remap['output_layers.0.' + k[len(f'output_layers.{i}.'):]] = t
                elif k.startswith(f'scalers.{i}.'):
                    remap['scalers.0.' + k[len(f'scalers.{i}.'):]] = t
                elif d and k.startswith(f'output_layers.{d[0]}.'):
                    remap['direct_head.' + k[len(f'output_layers.{d[0]}.'):]] = t
                elif d and k == f'scalers.{d[0]}.scale.weight':
                    remap['direct_head.scale'] = t
                elif not (k.startswith('output_layers.') or k.startswith('scalers.')):
                    remap[k] = t
            self._pack = WeightPack(remap, self.cutoff, device)
            self._pack_key = key
        return self._pack

    # ------------------------------------------------------------------ forward
    def forward(self, z, pos, cell, batch):
        """Network forward pass (reference models/newtonnet.py:74-104).

        z [N] int64 atomic numbers, pos [N,3], cell [B,3,3] (zeros = not periodic), batch [N] int64
        non-decreasing system index.  Returns a CustomOutputSet with z, pos, cell, batch, displacement,
        atom_node, force_node, edge_index and one attribute per entry of `output_properties`:
        energy [B], gradient_force [N,3], stress [B,3,3], virial [B,3,3], direct_force [N,3].
        """
        from newtonnet_b200.engine import get_engine
        props = list(self.output_properties)
        for key in props:
            if key not in _SUPPORTED:
                raise NotImplementedError(f"output '{key}' is outside the B200 energy/force/stress path")
        # training (reference train/trainer.py:303-313 differentiates the loss w.r.t. the parameters, also for energy-only
        # or direct_force-only heads, train/loss.py): outputs must carry a grad_fn -> autograd-composable path
        trainable = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if 'hessian' in props or trainable or any(getattr(layer, 'create_graph', False) for layer in self.output_layers):
            from newtonnet_b200.train import differentiable_forward
            return differentiable_forward(self, z, pos, cell, batch)
        if not pos.is_cuda:
            raise RuntimeError('newtonnet_b200: inputs must be CUDA tensors - there is no CPU fallback')
        derivative = [k for k in props if k not in ('energy', 'direct_force')]
        if derivative and props.index('energy') > min(props.index(k) for k in derivative):
            raise AttributeError("'CustomOutputSet' object has no attribute 'energy'")   # as the reference
        if self.embedding_layers.requires_dr and pos.is_leaf and not pos.requires_grad and pos.is_floating_point():
            pos.requires_grad = True      # side effect of models/newtonnet.py:150-152, kept for drop-in parity
        want_virial = 'stress' in props or 'virial' in props
        want_forces = 'gradient_force' in props or want_virial
        engine = get_engine(pos.device)
        pack = self._weight_pack(pos.device)
        res = engine.energy_forces(pack, z, pos, cell, batch, want_forces=want_forces, want_virial=want_virial,
                                   want_nodes=self.return_node_features, want_direct='direct_force' in props)
        dt = pos.dtype
        nl = res['_nl']
        displacement = torch.eye(3, dtype=dt, device=pos.device).repeat(cell.shape[0], 1, 1)
        generation = nl.generation
        outputs = CustomOutputSet(_lazy={'edge_index': lambda: nl.edge_index(generation)}, z=z, pos=pos, cell=cell,
                                  displacement=displacement, batch=batch)
        if self.return_node_features:
            outputs.atom_node = res['atom_node'].to(dt)
            outputs.force_node = res['force_node'].to(dt)
        if want_forces:
            outputs.pos_grad = -res['forces'].to(dt)
        if want_virial:
            outputs.displacement_grad = -res['virial'].to(dt)
        for key in props:
            if key == 'energy':
                value = res['energy'].to(dt)
            elif key == 'gradient_force':
                value = res['forces'].to(dt)
            elif key == 'virial':
                value = res['virial'].to(dt)
            elif key == 'direct_force':
                value = res['direct_force'].to(dt)
            else:
                value = res['stress'].to(dt)
            setattr(outputs, key, value)
        outputs.neighbor_list = nl
        return outputs

    def train(self, mode=True):
        """Training mode switches the derivative heads to create_graph (models/newtonnet.py:106-113)."""
        super().train(mode)
        for output_layer in self.output_layers:
            if isinstance(output_layer, DerivativeProperty):
                output_layer.create_graph = mode
        return self


class EmbeddingNet(nn.Module):
    """Atom / edge embedding parameters (reference models/newtonnet.py:116-137)."""

    def __init__(self, cutoff, n_features, n_basis):
        super().__init__()
        self.n_features = n_features
        self.node_embedding = nn.Embedding(118 + 1, n_features, padding_idx=0)
        self.edge_embedding = EdgeEmbedding(cutoff=cutoff, n_basis=n_basis)
        self.requires_dr = False


class InteractionNet(nn.Module):
    """Message-passing layer parameters (reference models/newtonnet.py:175-205)."""

    def __init__(self, n_features, n_basis, activation, layer_norm):
        super().__init__()
        self.n_features = n_features
        self.message_nodepart = nn.Sequential(
            nn.Linear(n_features, n_features), activation, nn.Linear(n_features, n_features))
        self.message_edgepart = nn.Linear(n_basis, n_features, bias=False)
        self.equiv_message1 = nn.Sequential(
            nn.Linear(n_features, n_features, bias=False), activation, nn.Linear(n_features, n_features, bias=False))
        self.equiv_message2 = nn.Sequential(
            nn.Linear(n_features, n_features, bias=False), activation, nn.Linear(n_features, n_features, bias=False))
        self.equiv_update = nn.Linear(n_features, n_features, bias=False)
        self.layer_norm = nn.LayerNorm(n_features) if layer_norm else None
