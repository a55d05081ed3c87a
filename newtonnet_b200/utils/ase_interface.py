"""ASE calculator - mirrors newtonnet/utils/ase_interface.py (MLAseCalculator).

Same constructor, `implemented_properties`, `calculate()` contract and result layout; the model behind it
is `newtonnet_b200.NewtonNet`, i.e. the CUDA library.  ASE itself is optional: when it is not installed
a minimal stand-in for `ase.calculators.calculator.Calculator` is used and any object exposing
get_atomic_numbers / get_positions(wrap=) / get_cell / get_pbc is accepted as `Atoms`.
"""
import numpy as np
import torch

from newtonnet_b200.compat import load_model as _load_checkpoint
from newtonnet_b200.layers.precision import get_precision_by_string
from newtonnet_b200.layers.scalers import get_scaler_by_string
from newtonnet_b200.models.output import (DerivativeProperty, SecondDerivativeProperty, get_aggregator_by_string,
                                          get_output_by_string)
from newtonnet_b200.utils.pretrained_models import download_checkpoint

__all__ = ['MLAseCalculator']

try:   # pragma: no cover - ase is not part of the build image
    from ase import Atoms as _Atoms
    from ase.calculators.calculator import Calculator as _Calculator
except ImportError:
    _Atoms = None

    class _Calculator:
        """Just enough of ase.calculators.calculator.Calculator for MD-style drivers and tests."""
        implemented_properties = []

        def __init__(self, **kwargs):
            self.atoms = None
            self.results = {}
            self.parameters = dict(kwargs)

        def calculate(self, atoms=None, properties=None, system_changes=None):
            if atoms is not None:
                self.atoms = atoms.copy() if hasattr(atoms, 'copy') and not isinstance(atoms, list) else atoms

        def get_property(self, name, atoms=None):
            self.calculate(atoms, [name], None)
            return self.results[name]

        def get_potential_energy(self, atoms=None):
            return self.get_property('energy', atoms)

        def get_forces(self, atoms=None):
            return self.get_property('forces', atoms)

        def get_stress(self, atoms=None):
            return self.get_property('stress', atoms)


def _is_single(atoms):
    if _Atoms is not None and isinstance(atoms, _Atoms):
        return True
    return hasattr(atoms, 'get_atomic_numbers')


class MLAseCalculator(_Calculator):
    implemented_properties = ['charges', 'bec', 'energy', 'free_energy', 'forces', 'hessian', 'stress']

    def __init__(self, model_path, properties: list = None, device: str = None, precision: str = 'float32',
                 **kwargs):
        """model_path: checkpoint path (reference pickle or state dict) or a NewtonNet module;
        properties: ASE property names (default: those of the model); device: CUDA device (default cuda);
        precision: 'float32' (kernel dtype; other precisions are cast at the boundary)."""
        _Calculator.__init__(self, **kwargs)
        self.device = torch.device('cuda' if device is None else device)
        if self.device.type != 'cuda':
            raise RuntimeError('newtonnet_b200 calculators run on CUDA devices only (no CPU fallback)')
        self.dtype = get_precision_by_string(precision)
        self.properties = properties
        self.model = self.load_model(model_path)
        self._pinned = {}
        self._resident = None        # device-resident step state of the last system shape (see _calculate_resident)

    # -------------------------------------------------------------- calculate (ase_interface.py:52-81)
    def calculate(self, atoms=None, properties=None, system_changes=None):
        _Calculator.calculate(self, atoms, self.properties, system_changes)
        if _is_single(atoms):
            atoms = [atoms]
        n_frames, n_atoms = len(atoms), len(atoms[0])
        host = self._host_arrays(atoms)
        if self._calculate_resident(host, n_frames, n_atoms):
            return
        z, pos, cell, batch = self._upload(host)
        pred = self.model(z, pos, cell, batch)
        for key in self.properties:
            if key in ('charges', 'bec'):
                raise NotImplementedError(f"property '{key}' is outside the B200 energy/force/stress path")
        if 'energy' in self.properties or 'free_energy' in self.properties:
            energy = pred.energy.cpu().detach().numpy()
            if 'energy' in self.properties:
                self.results['energy'] = energy.squeeze()
            if 'free_energy' in self.properties:
                self.results['free_energy'] = energy.squeeze()
        if 'forces' in self.properties:
            force = pred.gradient_force.cpu().detach().numpy()
            self.results['forces'] = force.reshape(n_frames, n_atoms, 3).squeeze()
        if 'hessian' in self.properties:
            hessian = pred.hessian.cpu().detach().numpy()
            self.results['hessian'] = hessian.reshape(n_frames, n_atoms, 3, n_atoms, 3).squeeze()
        if 'stress' in self.properties:
            stress = pred.stress.cpu().detach().numpy()
            self.results['stress'] = stress[:, [0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]].squeeze()   # Voigt
        del pred

    # -------------------------------------------------------------- model load (ase_interface.py:83-129)
    def load_model(self, model):
        if isinstance(model, str) and model in ['ani1', 'ani1x', 't1x']:
            model = download_checkpoint(model)
        model = _load_checkpoint(model, map_location=self.device)
        model.return_node_features = False
        if self.properties is None:
            names = {'charge': 'charges', 'energy': 'energy', 'gradient_force': 'forces'}
            self.properties = [names.get(key) for key in model.output_properties]
        else:
            model.output_properties = list(model.output_properties)
            keys_to_keep = ['charge', 'energy']
            names = {'charges': 'charge', 'bec': 'bec', 'energy': 'energy', 'free_energy': 'energy',
                     'forces': 'gradient_force', 'stress': 'stress', 'hessian': 'hessian'}
            for key in self.properties:
                key = names.get(key)
                keys_to_keep.append(key)
                if key in model.output_properties:
                    continue
                model.output_properties.append(key)
                model.output_layers.append(get_output_by_string(key))
                model.scalers.append(get_scaler_by_string(key))
                model.aggregators.append(get_aggregator_by_string(key))
            ids_to_remove = [i for i, key in enumerate(model.output_properties) if key not in keys_to_keep]
            for i in reversed(ids_to_remove):
                model.output_properties.pop(i)
                model.output_layers.pop(i)
                model.scalers.pop(i)
                model.aggregators.pop(i)
        model.to(self.dtype)
        model.eval()
        model.embedding_layers.requires_dr = any(isinstance(layer, DerivativeProperty) for layer in model.output_layers)
        if any(isinstance(layer, SecondDerivativeProperty) for layer in model.output_layers):
            for layer in model.output_layers:
                if isinstance(layer, DerivativeProperty):
                    layer.create_graph = True
        return model

    # -------------------------------------------------------------- input staging (ase_interface.py:131-142)
    def _staging(self, name, shape, dtype):
        buf = self._pinned.get(name)
        if buf is None or buf.shape != shape or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[name] = buf
        return buf

    def _param_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def _host_arrays(self, atoms_list):
        """Atoms -> numpy (z, pos, cell, batch): wrapped positions, zero rows for non-periodic directions
        (ase_interface.py:134-138)."""
        zs, ps, cs, bs = [], [], [], []
        for b, atoms in enumerate(atoms_list):
            z = np.asarray(atoms.get_atomic_numbers(), dtype=np.int64)
            pos = np.asarray(atoms.get_positions(wrap=True), dtype=np.float64)
            cell = np.array(getattr(atoms.get_cell(), 'array', atoms.get_cell()), dtype=np.float64).reshape(3, 3)
            pbc = np.asarray(atoms.get_pbc(), dtype=bool).reshape(3)
            cell = cell.copy()
            cell[~pbc] = 0.0
            zs.append(z); ps.append(pos); cs.append(cell); bs.append(np.full(len(z), b, dtype=np.int64))
        np_dtype = {torch.float32: np.float32, torch.float64: np.float64, torch.float16: np.float16}[self.dtype]
        return {'z': np.concatenate(zs), 'pos': np.concatenate(ps).astype(np_dtype),
                'cell': np.stack(cs).astype(np_dtype), 'batch': np.concatenate(bs)}

    def _upload(self, host):
        out = []
        for name in ('z', 'pos', 'cell', 'batch'):
            src = torch.from_numpy(host[name])
            stage = self._staging(name, src.shape, src.dtype)
            stage.copy_(src)
            out.append(stage.to(self.device, non_blocking=True))
        return tuple(out)

    def format_data(self, atoms_list):
        """Atoms -> (z, pos, cell, batch) on the device; one pinned staging copy + async H2D per tensor."""
        return self._upload(self._host_arrays(atoms_list))

    # -------------------------------------------------------------- device-resident MD-step path
    def _calculate_resident(self, host, n_frames, n_atoms):
        """Steady-state path of an MD driver (reference: ase_interface.py:52-81 re-uploads z / cell / batch, launches the
        whole model from Python and reads every result back with its own synchronisation, every step).  From the second
        call with the same system shape on, atomic numbers, cell and batch stay resident on the device (re-sent only when
        they change), the positions go from a pinned buffer straight into the static input of ONE CUDA graph (neighbour
        rebuild + evaluation), and status, energy, forces and stress come back as asynchronous copies into pinned memory
        behind a single stream synchronisation.  Returns False when the general path has to run."""
        from newtonnet_b200 import _lib as L
        from newtonnet_b200.engine import GraphedStep, get_engine
        props = self.properties
        if self.dtype != torch.float32 or any(k not in ('energy', 'free_energy', 'forces', 'stress') for k in props):
            return False
        mp = list(self.model.output_properties)
        if any(k not in ('energy', 'gradient_force', 'stress', 'virial') for k in mp) or 'energy' not in mp:
            return False
        if any(getattr(layer, 'create_graph', False) for layer in self.model.output_layers) or self.model.training:
            return False
        engine = get_engine(self.device)
        if not engine.use_cuda_graphs:
            return False
        want_virial = 'stress' in mp or 'virial' in mp
        want_forces = 'gradient_force' in mp or want_virial
        N, B = host['pos'].shape[0], host['cell'].shape[0]
        key = (N, B, want_forces, want_virial, tuple(props))
        st = self._resident
        if st is None or st['key'] != key:
            if st is not None and st.get('pending_key') == key and engine._nl is not None and engine._nl.n_atoms == N:
                # second call with this shape: capacities are known from the general path's call -> capture the step
                pack = self.model._weight_pack(self.device)
                z, pos, cell, batch = self._upload(host)
                step = GraphedStep(engine, pack, z, pos, cell, batch, want_forces, want_virial, False, engine._nl.cap_edges)
                pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
                self._resident = st = {
                    'key': key, 'step': step, 'pack': pack, 'sig': self._param_signature(),
                    'host': {k: host[k].copy() for k in ('z', 'cell', 'batch')},
                    'pos_pin': pin((N, 3), torch.float32), 'status_pin': pin((L.NN_STATUS_WORDS,), torch.int32),
                    'e_pin': pin((B,), torch.float32), 'f_pin': pin((N, 3), torch.float32) if want_forces else None,
                    's_pin': pin((B, 3, 3), torch.float32) if want_virial else None}
            else:
                self._resident = {'key': None, 'pending_key': key}
                return False
        step = st['step']
        if st['sig'] != self._param_signature():                          # parameters changed: recapture via the general path
            self._resident = None
            return False
        for name, dst in (('z', step.z), ('cell', step.cell), ('batch', step.batch)):
            if not np.array_equal(host[name], st['host'][name]):
                dst.copy_(torch.from_numpy(host[name]).to(dst.dtype).reshape(dst.shape))
                st['host'][name] = host[name].copy()
        st['pos_pin'].copy_(torch.from_numpy(host['pos']))
        step.pos.copy_(st['pos_pin'], non_blocking=True)
        step.nl.n_edges = None
        step.nl.generation += 1
        step.graph.replay()
        st['status_pin'].copy_(step.nl.status, non_blocking=True)
        st['e_pin'].copy_(step.out['energy'], non_blocking=True)
        if want_forces:
            st['f_pin'].copy_(step.out['forces'], non_blocking=True)
        if want_virial:
            st['s_pin'].copy_(step.out['stress'], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()             # the one synchronisation of the step
        status = st['status_pin'].tolist()
        if status[L.ST_EDGE_OVERFLOW] or status[L.ST_ROW_OVERFLOW] or status[L.ST_BATCH_UNSORTED] or status[L.ST_SINGULAR_CELL]:
            self._resident = None            # capacity outgrown or bad input: the general path regrows / raises
            return False
        energy = st['e_pin'].numpy().copy()
        if 'energy' in props:
            self.results['energy'] = energy.squeeze()
        if 'free_energy' in props:
            self.results['free_energy'] = energy.squeeze()
        if 'forces' in props:
            self.results['forces'] = st['f_pin'].numpy().copy().reshape(n_frames, n_atoms, 3).squeeze()
        if 'stress' in props:
            stress = st['s_pin'].numpy().copy()
            self.results['stress'] = stress[:, [0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]].squeeze()   # Voigt
        return True
