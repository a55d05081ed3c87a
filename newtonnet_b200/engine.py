"""Host driver of the CUDA path: neighbour list, weight pack, one-call evaluation.

PyTorch is used for device memory and streams only; all arithmetic happens in
libnewtonnet_b200.so (include/newtonnet_b200.h).
"""
import ctypes as C
import os

import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'newtonnet_b200: `{name}` must be a CUDA tensor - this package has no CPU fallback '
                           f'(use the reference implementation or oracle/ for CPU).')


class WeightPack:
    """fp32 device copies of the reference parameters in both orientations (nn_weights)."""

    def __init__(self, state, cutoff, device):
        """`state`: mapping with the reference's parameter names -> tensors."""
        self.keep = []
        dev = device

        def f32(t):
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            self.keep.append(t)
            return t

        lib = L.load()
        stream = torch.cuda.current_stream(dev).cuda_stream

        def image(b):
            img = torch.empty(L.NN_B_IMAGE_FLOATS, dtype=torch.float32, device=dev)
            self.keep.append(img)
            L.check(lib.nn_gemm128_prepare_b(b.data_ptr(), img.data_ptr(), stream), 'nn_gemm128_prepare_b')
            return img.data_ptr()

        def mat(dst, t):
            """Both orientations of a [128,128] weight plus their tensor-core operand images."""
            w = f32(t)
            wt = f32(t.detach().t())
            dst.w, dst.wt = w.data_ptr(), wt.data_ptr()
            dst.w_img, dst.wt_img = image(w), image(wt)

        def both(t):
            w = f32(t)
            wt = f32(t.detach().t())
            return w.data_ptr(), wt.data_ptr()

        w = L.Weights()
        n_layers = 1 + max(int(k.split('.')[1]) for k in state if k.startswith('interaction_layers.'))
        if n_layers > L.NN_MAX_LAYERS:
            raise ValueError(f'n_interactions={n_layers} > {L.NN_MAX_LAYERS}')
        emb = state['embedding_layers.node_embedding.weight']
        if emb.shape[1] != L.NN_F:
            raise NotImplementedError(f'kernels are specialised for n_features={L.NN_F}, got {emb.shape[1]}')
        freq = state['embedding_layers.edge_embedding.embedding.frequencies']
        if freq.numel() != L.NN_NB:
            raise NotImplementedError(f'kernels are specialised for n_basis={L.NN_NB}, got {freq.numel()}')
        w.n_layers = n_layers
        w.cutoff = float(cutoff)
        w.embedding = f32(emb).data_ptr()
        w.frequencies = f32(freq).data_ptr()
        for l in range(n_layers):
            k = f'interaction_layers.{l}.'
            lw = w.layer[l]
            if (k + 'layer_norm.weight') in state:
                lw.ln_gamma = f32(state[k + 'layer_norm.weight']).data_ptr()
                lw.ln_beta = f32(state[k + 'layer_norm.bias']).data_ptr()
            mat(lw.W1, state[k + 'message_nodepart.0.weight'])
            lw.b1 = f32(state[k + 'message_nodepart.0.bias']).data_ptr()
            mat(lw.W2, state[k + 'message_nodepart.2.weight'])
            lw.b2 = f32(state[k + 'message_nodepart.2.bias']).data_ptr()
            lw.We, lw.Wet = both(state[k + 'message_edgepart.weight'])
            we_img = torch.empty(L.NN_WE_IMAGE_FLOATS, dtype=torch.float32, device=dev)
            self.keep.append(we_img)
            L.check(lib.nn_message_prepare_b(lw.We, we_img.data_ptr(), stream), 'nn_message_prepare_b')
            lw.We_img = we_img.data_ptr()
            mat(lw.U1, state[k + 'equiv_message1.0.weight'])
            mat(lw.U2, state[k + 'equiv_message1.2.weight'])
            mat(lw.V1, state[k + 'equiv_message2.0.weight'])
            mat(lw.V2, state[k + 'equiv_message2.2.weight'])
            mat(lw.Wu, state[k + 'equiv_update.weight'])
        h = 'output_layers.0.layers.'
        hk = [k for k in state if k.endswith('layers.0.weight') and k.startswith('output_layers.')]
        if hk:
            h = hk[0][:-len('0.weight')]
        idx = h.split('.')[1]
        mat(w.H1, state[h + '0.weight'])
        w.hb1 = f32(state[h + '0.bias']).data_ptr()
        mat(w.H2, state[h + '2.weight'])
        w.hb2 = f32(state[h + '2.bias']).data_ptr()
        w.w3 = f32(state[h + '4.weight'].reshape(-1)).data_ptr()
        w.hb3 = f32(state[h + '4.bias'].reshape(-1)).data_ptr()
        w.scale = f32(state[f'scalers.{idx}.scale.weight'].reshape(-1)).data_ptr()
        w.shift = f32(state[f'scalers.{idx}.shift.weight'].reshape(-1)).data_ptr()
        self.has_direct_force = 'direct_head.layers.0.weight' in state
        if self.has_direct_force:
            mat(w.D1, state['direct_head.layers.0.weight']); w.db1 = f32(state['direct_head.layers.0.bias']).data_ptr()
            mat(w.D2, state['direct_head.layers.2.weight']); w.db2 = f32(state['direct_head.layers.2.bias']).data_ptr()
            mat(w.D3, state['direct_head.layers.4.weight']); w.db3 = f32(state['direct_head.layers.4.bias']).data_ptr()
            w.dscale = f32(state['direct_head.scale'].reshape(-1)).data_ptr()
        self.struct = w
        self.n_layers = n_layers
        self.cutoff = float(cutoff)


class NeighborList:
    """Destination-sorted CSR + undirected pair list on the device (nn_nbr)."""

    def __init__(self, engine, pos, cell, batch, cap_edges, cap_pairs=None):
        """cap_pairs: capacity of the pair table; default cap_edges / 2 (complete lists have P = E / 2 - lists with empty
        ghost rows, domain decomposition, have up to P = E pairs and pass it explicitly)."""
        N, B = pos.shape[0], cell.shape[0]
        dev = pos.device
        i32 = dict(dtype=torch.int32, device=dev)
        self.pos, self.cell, self.batch = pos, cell, batch
        self.n_atoms, self.n_systems = N, B
        self.cap_edges = int(cap_edges)
        self.cap_pairs = (self.cap_edges + 1) // 2 if cap_pairs is None else int(cap_pairs)
        self.sys_ptr = torch.empty(B + 1, **i32)
        self.row_ptr = torch.empty(N + 1, **i32)
        self.pair_ptr = torch.empty(N + 1, **i32)
        self.col = torch.empty(max(self.cap_edges, 1), **i32)
        self.edge_pair = torch.empty(max(self.cap_edges, 1), **i32)
        self.pair_i = torch.empty(max(self.cap_pairs, 1), **i32)
        self.pair_j = torch.empty(max(self.cap_pairs, 1), **i32)
        self.pair_disp = torch.empty(max(self.cap_pairs, 1), 3, dtype=torch.float32, device=dev)
        self.status = torch.zeros(L.NN_STATUS_WORDS, **i32)
        ws_bytes = engine.lib.nn_nbr_workspace_bytes(N, B)
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        s = L.Nbr()
        s.n_atoms, s.n_systems = N, B
        s.cap_edges, s.cap_pairs, s.cap_cells = self.cap_edges, self.cap_pairs, 2 * N + B
        s.pos, s.cell, s.batch = pos.data_ptr(), cell.data_ptr(), batch.data_ptr()
        s.sys_ptr, s.row_ptr, s.col = self.sys_ptr.data_ptr(), self.row_ptr.data_ptr(), self.col.data_ptr()
        s.edge_pair, s.pair_ptr = self.edge_pair.data_ptr(), self.pair_ptr.data_ptr()
        s.pair_i, s.pair_j, s.pair_disp = self.pair_i.data_ptr(), self.pair_j.data_ptr(), self.pair_disp.data_ptr()
        s.status = self.status.data_ptr()
        s.workspace, s.workspace_bytes = self.workspace.data_ptr(), ws_bytes
        self.struct = s
        self.n_edges = None   # known on the host after `check()`
        self.generation = 0   # bumped every time the list is rebuilt in place

    def rebind(self, pos, cell, batch):
        """Point the list at new input tensors of the same shapes (next MD step)."""
        self.pos, self.cell, self.batch = pos, cell, batch
        self.struct.pos, self.struct.cell, self.struct.batch = pos.data_ptr(), cell.data_ptr(), batch.data_ptr()

    def check(self):
        """Read the status words (one small D2H copy, synchronises the stream)."""
        st = self.status.cpu().tolist()
        if st[L.ST_BATCH_UNSORTED] == 1:
            raise ValueError('batch must be non-decreasing with values in [0, n_systems)')
        if st[L.ST_BATCH_UNSORTED] == 3:
            raise ValueError('atomic numbers must be in [0, 118]')
        if st[L.ST_BATCH_UNSORTED] == 2:
            raise RuntimeError('internal error: asymmetric edge set')
        if st[L.ST_SINGULAR_CELL]:
            # the reference raises torch._C._LinAlgError from linalg.solve (representations.py:92)
            raise RuntimeError('singular cell in a periodic batch (mixed periodic / non-periodic systems or '
                               'partially periodic cells are not supported, as in the reference)')
        if st[L.ST_ROW_OVERFLOW]:
            raise RuntimeError(f'an atom has {st[L.ST_ROW_OVERFLOW]} neighbours (> {512})')
        self.n_edges = st[L.ST_N_EDGES]
        return st

    def edge_index(self, generation=None):
        """[2,E] int64, reference order (i-major, j ascending)."""
        if generation is not None and generation != self.generation:
            raise RuntimeError('edge_index of an earlier forward() was requested after the neighbour list had been '
                               'rebuilt in place by a later call; read it before the next forward()')
        if self.n_edges is None:
            self.check()
        out = torch.empty(2, self.n_edges, dtype=torch.int64, device=self.pos.device)
        if self.n_edges:
            lib = L.load()
            L.check(lib.nn_nbr_edge_index(C.byref(self.struct), out.data_ptr(), self.n_edges, _stream()), 'nn_nbr_edge_index')
        return out


class Engine:
    """Per-device driver.  Keeps capacities and workspaces between calls so a steady-state evaluation
    allocates nothing and synchronises once (the status read that accompanies the results)."""

    HEADROOM = 1.08

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('newtonnet_b200 runs on CUDA devices only (no CPU fallback)')
        self.lib = L.load()
        self._nl = None
        self._ws = None
        self.launches = 0
        self.use_cuda_graphs = os.environ.get('NN_CUDA_GRAPHS', '1') != '0'
        self._graphs = {}
        self._last_key = None

    # ------------------------------------------------------------------ neighbour list
    def neighbor_list(self, pos, cell, batch, cutoff):
        for t, n in ((pos, 'pos'), (cell, 'cell'), (batch, 'batch')):
            _require_cuda(t, n)
        pos = pos.detach().to(torch.float32).contiguous()
        cell = cell.detach().to(torch.float32).contiguous().reshape(-1, 3, 3)
        batch = batch.to(torch.int64).contiguous()
        N, B = pos.shape[0], cell.shape[0]
        nl = self._nl
        reuse = nl is not None and nl.n_atoms == N and nl.n_systems == B and nl.pos.device == pos.device
        s = _stream()
        if reuse:
            nl.rebind(pos, cell, batch)
            nl.n_edges = None
            nl.generation += 1
            L.check(self.lib.nn_nbr_count(C.byref(nl.struct), cutoff, s), 'nn_nbr_count')
            L.check(self.lib.nn_nbr_fill(C.byref(nl.struct), cutoff, s), 'nn_nbr_fill')
            return nl   # overflow (if any) is detected by the caller's status check -> `grow`
        probe = NeighborList(self, pos, cell, batch, cap_edges=0)
        L.check(self.lib.nn_nbr_count(C.byref(probe.struct), cutoff, s), 'nn_nbr_count')
        st = probe.check()
        return self._build_with_capacity(pos, cell, batch, cutoff, st[L.ST_N_EDGES])

    def _build_with_capacity(self, pos, cell, batch, cutoff, n_edges):
        cap = int(n_edges * self.HEADROOM) + 64
        cap += cap % 2
        nl = NeighborList(self, pos, cell, batch, cap_edges=cap)
        s = _stream()
        L.check(self.lib.nn_nbr_count(C.byref(nl.struct), cutoff, s), 'nn_nbr_count')
        L.check(self.lib.nn_nbr_fill(C.byref(nl.struct), cutoff, s), 'nn_nbr_fill')
        self._nl = nl
        return nl

    def grow(self, nl, cutoff, needed):
        return self._build_with_capacity(nl.pos, nl.cell, nl.batch, cutoff, needed)

    def checked_neighbor_list(self, pos, cell, batch, cutoff):
        """neighbor_list + status read; a cached list whose capacity the new positions outgrow is rebuilt larger
        (the sync-free `neighbor_list` leaves that to the caller).  Returns a list that is complete."""
        nl = self.neighbor_list(pos, cell, batch, cutoff)
        st = nl.check()
        if st[L.ST_EDGE_OVERFLOW]:
            nl = self.grow(nl, cutoff, max(st[L.ST_EDGE_OVERFLOW], st[L.ST_N_EDGES]))
            st = nl.check()
            if st[L.ST_EDGE_OVERFLOW]:
                raise RuntimeError('neighbour list capacity overflow after regrow')
        return nl

    # ------------------------------------------------------------------ evaluation
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != self.device:
            self._ws = torch.empty(int(nbytes * 1.05) + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    def evaluate(self, nl, weights, z, want_forces=True, want_virial=False, want_nodes=False, want_direct=False):
        _require_cuda(z, 'z')
        z = z.to(torch.int64).contiguous()
        N, B = nl.n_atoms, nl.n_systems
        dev = nl.pos.device
        f32 = dict(dtype=torch.float32, device=dev)
        out = {'energy': torch.empty(B, **f32)}
        bwd = want_forces or want_virial
        if bwd:
            out['forces'] = torch.empty(N, 3, **f32)
        if want_virial:
            out['virial'] = torch.empty(B, 3, 3, **f32)
            out['stress'] = torch.empty(B, 3, 3, **f32)
        if want_nodes:
            out['atom_node'] = torch.empty(N, L.NN_F, **f32)
            out['force_node'] = torch.empty(N, 3, L.NN_F, **f32)
        if want_direct:
            out['direct_force'] = torch.empty(N, 3, **f32)
        nbytes = self.lib.nn_eval_workspace_bytes(N, B, nl.cap_pairs, weights.n_layers, int(bwd))
        ws = self._workspace(nbytes)
        a = L.EvalArgs()
        a.nbr = C.pointer(nl.struct)
        a.w = C.pointer(weights.struct)
        a.z = z.data_ptr()
        a.want_forces, a.want_virial = int(bwd), int(want_virial)
        a.energy = out['energy'].data_ptr()
        a.forces = L.ptr(out.get('forces'))
        a.virial = L.ptr(out.get('virial'))
        a.stress = L.ptr(out.get('stress'))
        a.atom_node = L.ptr(out.get('atom_node'))
        a.force_node = L.ptr(out.get('force_node'))
        a.direct_force = L.ptr(out.get('direct_force'))
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(self.lib.nn_eval(C.byref(a), _stream()), 'nn_eval')
        out['_z'] = z   # keep alive until the stream has consumed it
        return out

    def energy_forces(self, weights, z, pos, cell, batch, want_forces=True, want_virial=False, want_nodes=False,
                      want_direct=False):
        """Neighbour list + evaluation + status check, regrowing capacities when needed.  From the second
        call with the same shapes on, the whole step (neighbour rebuild + ~80 kernels) is replayed as one
        CUDA graph."""
        if self.use_cuda_graphs:
            key = (pos.shape[0], cell.reshape(-1, 9).shape[0], bool(want_forces), bool(want_virial), bool(want_nodes),
                   bool(want_direct), id(weights), str(pos.device))
            g = self._graphs.get(key)
            if g is not None:
                out = g.replay(z, pos, cell, batch)
                if out is not None:
                    return out
                del self._graphs[key]                         # capacity overflow: fall through, recapture later
            elif self._last_key == key and self._nl is not None and self._nl.n_atoms == pos.shape[0]:
                if len(self._graphs) >= 4:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = GraphedStep(self, weights, z, pos, cell, batch, want_forces, want_virial, want_nodes,
                                                self._nl.cap_edges, want_direct)
                out = self._graphs[key].replay(z, pos, cell, batch)
                if out is not None:
                    return out
                del self._graphs[key]
            self._last_key = key
        nl = self.neighbor_list(pos, cell, batch, weights.cutoff)
        out = self.evaluate(nl, weights, z, want_forces, want_virial, want_nodes, want_direct)
        st = nl.check()
        if st[L.ST_EDGE_OVERFLOW]:
            nl = self.grow(nl, weights.cutoff, max(st[L.ST_EDGE_OVERFLOW], st[L.ST_N_EDGES]))
            out = self.evaluate(nl, weights, z, want_forces, want_virial, want_nodes, want_direct)
            st = nl.check()
            if st[L.ST_EDGE_OVERFLOW]:
                raise RuntimeError('neighbour list capacity overflow after regrow')
        out['_nl'] = nl
        return out


class GraphedStep:
    """One evaluation (neighbour rebuild + nn_eval) captured as a CUDA graph over static buffers."""

    def __init__(self, engine, weights, z, pos, cell, batch, want_forces, want_virial, want_nodes, cap_edges,
                 want_direct=False):
        dev = pos.device
        self.engine, self.weights = engine, weights
        self.z = torch.empty(z.shape, dtype=torch.int64, device=dev)
        self.pos = torch.empty(pos.shape, dtype=torch.float32, device=dev)
        self.cell = torch.empty(cell.reshape(-1, 3, 3).shape, dtype=torch.float32, device=dev)
        self.batch = torch.empty(batch.shape, dtype=torch.int64, device=dev)
        self._copy_in(z, pos, cell, batch)
        self.nl = NeighborList(engine, self.pos, self.cell, self.batch, cap_edges=cap_edges)
        N, B = self.nl.n_atoms, self.nl.n_systems
        f32 = dict(dtype=torch.float32, device=dev)
        bwd = want_forces or want_virial
        self.out = {'energy': torch.empty(B, **f32)}
        if bwd:
            self.out['forces'] = torch.empty(N, 3, **f32)
        if want_virial:
            self.out['virial'] = torch.empty(B, 3, 3, **f32)
            self.out['stress'] = torch.empty(B, 3, 3, **f32)
        if want_nodes:
            self.out['atom_node'] = torch.empty(N, L.NN_F, **f32)
            self.out['force_node'] = torch.empty(N, 3, L.NN_F, **f32)
        if want_direct:
            self.out['direct_force'] = torch.empty(N, 3, **f32)
        nbytes = engine.lib.nn_eval_workspace_bytes(N, B, self.nl.cap_pairs, weights.n_layers, int(bwd))
        self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        a = L.EvalArgs()
        a.nbr, a.w, a.z = C.pointer(self.nl.struct), C.pointer(weights.struct), self.z.data_ptr()
        a.want_forces, a.want_virial = int(bwd), int(want_virial)
        a.energy = self.out['energy'].data_ptr()
        a.forces, a.virial, a.stress = L.ptr(self.out.get('forces')), L.ptr(self.out.get('virial')), L.ptr(self.out.get('stress'))
        a.atom_node, a.force_node = L.ptr(self.out.get('atom_node')), L.ptr(self.out.get('force_node'))
        a.direct_force = L.ptr(self.out.get('direct_force'))
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        self.args = a
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._launch()

    def _launch(self):
        lib, s = self.engine.lib, _stream()
        L.check(lib.nn_nbr_count(C.byref(self.nl.struct), self.weights.cutoff, s), 'nn_nbr_count')
        L.check(lib.nn_nbr_fill(C.byref(self.nl.struct), self.weights.cutoff, s), 'nn_nbr_fill')
        L.check(lib.nn_eval(C.byref(self.args), s), 'nn_eval')

    def _copy_in(self, z, pos, cell, batch):
        self.z.copy_(z, non_blocking=True)
        self.pos.copy_(pos.detach(), non_blocking=True)
        self.cell.copy_(cell.detach().reshape(-1, 3, 3), non_blocking=True)
        self.batch.copy_(batch, non_blocking=True)

    def replay(self, z, pos, cell, batch):
        """Returns the result dict, or None when the captured capacities overflowed (caller regrows)."""
        for t, n in ((pos, 'pos'), (cell, 'cell'), (batch, 'batch'), (z, 'z')):
            _require_cuda(t, n)
        self._copy_in(z, pos, cell, batch)
        self.nl.n_edges = None
        self.nl.generation += 1
        self.graph.replay()
        out = {k: v.clone() for k, v in self.out.items()}
        st = self.nl.check()
        if st[L.ST_EDGE_OVERFLOW]:
            return None
        out['_nl'] = self.nl
        return out


_engines = {}


def get_engine(device):
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    key = str(device)
    if key not in _engines:
        _engines[key] = Engine(device)
    return _engines[key]
