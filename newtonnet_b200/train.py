"""Differentiable (create_graph) evaluation and the data-parallel training step - SURVEY.md section 8a row T.

The inference path (csrc/eval.cu) uses a hand-derived reverse sweep and is not differentiable with respect
to the parameters.  Training needs d(loss)/d(parameters) THROUGH the forces (reference
models/newtonnet.py:106-113 sets create_graph=True; train/trainer.py:303-313 calls loss.backward()).  Here
the forward is composed from a closed set of primitives whose derivatives are again those primitives, so
autograd builds the double backward:

    X @ B         nn_gemm128 (tcgen05 3xTF32)     dX = dY @ B^T  (same op)      dB = X^T dY  (nn_gemm128_tn)
    X^T Y         nn_gemm128_tn                   dX = Y @ G^T                  dY = X @ G   (nn_gemm128)
    rows[idx]     nn_halo_pack (gather)           d rows = segment sum
    segment sum   nn_segment_sum (deterministic)  d src = rows[idx]

i.e. every 128-wide contraction (98 % of the FLOPs) and every gather / scatter of feature rows runs in the
CUDA library; element-wise glue (SiLU, products, the radial basis) is left to autograd's own element-wise
kernels in this round.  Edges come from the same cell-list kernel as inference (directed edges in the
reference's order).  NewtonNet.forward routes here when a derivative head has create_graph=True.
"""
import ctypes as C

import torch
import torch.distributed as dist
import torch.nn.functional as Fn

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------- raw kernel calls
# Operand images of the model's weights, valid for ONE training step: every nn.Linear weight is multiplied in both
# orientations several times per step (forward, backward, double backward).  Only views of leaf parameters are cached
# (temporaries can reuse an address), keyed by storage address, version and strides; differentiable_forward clears the
# cache when a step begins, so a CUDA-graph capture always records the prepare kernels of its own step.
_image_cache = {}


def _operand_image(B, Bc):
    lib = L.load()
    base = B._base if B._base is not None else B
    key = None
    if base.is_leaf and isinstance(base, torch.nn.Parameter):
        key = (B.data_ptr(), base._version, tuple(B.shape), tuple(B.stride()))
        img = _image_cache.get(key)
        if img is not None:
            return img
    img = torch.empty(L.NN_B_IMAGE_FLOATS, dtype=torch.float32, device=Bc.device)
    L.check(lib.nn_gemm128_prepare_b(Bc.data_ptr(), img.data_ptr(), _stream()), 'nn_gemm128_prepare_b')
    if key is not None:
        _image_cache[key] = img
    return img


def _gemm_raw(X, B, B_orig=None):
    lib = L.load()
    M = X.shape[0]
    Y = torch.empty(M, 128, dtype=torch.float32, device=X.device)
    if M == 0:
        return Y
    img = _operand_image(B if B_orig is None else B_orig, B)
    a = L.GemmArgs()
    a.X, a.B, a.B_img, a.Y, a.m = X.data_ptr(), B.data_ptr(), img.data_ptr(), Y.data_ptr(), M
    L.check(lib.nn_gemm128(C.byref(a), _stream()), 'nn_gemm128')
    return Y


def _gemm_tn_raw(X, Y):
    lib = L.load()
    M = X.shape[0]
    out = torch.empty(128, 128, dtype=torch.float32, device=X.device)
    ws = torch.empty(max(lib.nn_gemm128_tn_workspace_bytes(M), 4), dtype=torch.uint8, device=X.device)
    L.check(lib.nn_gemm128_tn(X.data_ptr(), Y.data_ptr(), M, out.data_ptr(), ws.data_ptr(), _stream()), 'nn_gemm128_tn')
    return out


def _c(t):
    return t.contiguous()


class Gemm(torch.autograd.Function):
    """Y[M,128] = X[M,128] @ B[128,128]."""

    @staticmethod
    def forward(ctx, X, B):
        # save the inputs themselves: a contiguous copy made here would be cut off from the graph, and
        # the double backward needs d(dX)/dB through them
        ctx.save_for_backward(X, B)
        return _gemm_raw(_c(X), _c(B), B)

    @staticmethod
    def backward(ctx, dY):
        X, B = ctx.saved_tensors
        dX = Gemm.apply(dY, B.t()) if ctx.needs_input_grad[0] else None
        dB = GemmTN.apply(X, dY) if ctx.needs_input_grad[1] else None
        return dX, dB


class GemmTN(torch.autograd.Function):
    """G[128,128] = X[M,128]^T @ Y[M,128]."""

    @staticmethod
    def forward(ctx, X, Y):
        ctx.save_for_backward(X, Y)
        return _gemm_tn_raw(_c(X), _c(Y))

    @staticmethod
    def backward(ctx, G):
        X, Y = ctx.saved_tensors
        dX = Gemm.apply(Y, G.t()) if ctx.needs_input_grad[0] else None
        dY = Gemm.apply(X, G) if ctx.needs_input_grad[1] else None
        return dX, dY


class Segments:
    """Index structure of one gather / segment-sum pair: idx[k] = target row of source row k."""

    def __init__(self, idx, n_rows, grouped=None, valid=None):
        """grouped: True = idx is non-decreasing, False = it is not, None = look (one host synchronisation; not
        allowed while a CUDA graph is being captured).  valid: mask of the real rows of a padded list; padding must
        come last - it is gathered (harmless) but belongs to no segment, so sums neither see it nor pay for it."""
        self.idx = idx.to(torch.int32).contiguous()
        self.n_rows = int(n_rows)
        ones = torch.ones(idx.shape[0], dtype=torch.int64, device=idx.device) if valid is None else valid.to(torch.int64)
        counts = torch.zeros(n_rows, dtype=torch.int64, device=idx.device).index_add_(0, idx.long(), ones)   # no host sync
        self.row_ptr = torch.zeros(n_rows + 1, dtype=torch.int32, device=idx.device)
        self.row_ptr[1:] = torch.cumsum(counts, 0)
        if grouped is None:
            grouped = bool((idx[1:] >= idx[:-1]).all()) if idx.numel() > 1 else True
        if grouped:
            self.perm = None                      # already grouped (destination-sorted edges)
        else:
            self.perm = torch.sort(idx, stable=True).indices.to(torch.int32).contiguous()


class Gather(torch.autograd.Function):
    """out[k,:] = rows[idx[k],:]  (width % 4 == 0)."""

    @staticmethod
    def forward(ctx, rows, seg):
        rows = _c(rows)
        ctx.seg = seg
        n, width = seg.idx.shape[0], rows.shape[1]
        out = torch.empty(n, width, dtype=torch.float32, device=rows.device)
        if n:
            L.check(L.load().nn_halo_pack(rows.data_ptr(), seg.idx.data_ptr(), n, width, out.data_ptr(), _stream()),
                    'nn_halo_pack')
        return out

    @staticmethod
    def backward(ctx, d_out):
        return SegmentSum.apply(d_out, ctx.seg), None


class SegmentSum(torch.autograd.Function):
    """out[i,:] = sum_{k: idx[k] = i} src[k,:], fixed summation order."""

    @staticmethod
    def forward(ctx, src, seg):
        src = _c(src)
        ctx.seg = seg
        width = src.shape[1]
        out = torch.zeros(seg.n_rows, width, dtype=torch.float32, device=src.device)
        if src.shape[0]:
            L.check(L.load().nn_segment_sum(src.data_ptr(), L.ptr(seg.perm), seg.row_ptr.data_ptr(), seg.n_rows, width,
                                            out.data_ptr(), _stream()), 'nn_segment_sum')
        return out

    @staticmethod
    def backward(ctx, d_out):
        return Gather.apply(d_out, ctx.seg), None


def linear(x, weight, bias=None):
    """x @ weight^T (+ bias) through the tensor-core GEMM."""
    y = Gemm.apply(x, weight.t())
    return y if bias is None else y + bias


# ----------------------------------------------------------------------------- forward (training mode)
def _envelope(x):
    # 1 - 55x^9 + 99x^10 - 45x^11 in the factored form used by csrc/pair_ops.cu
    p = torch.zeros_like(x) + 45.0
    for c in (36.0, 28.0, 21.0, 15.0, 10.0, 6.0, 3.0, 1.0):
        p = p * x + c
    return (1.0 - x) ** 3 * p


def _edges(nl, pos, N, static):
    """Directed edges (reference order) and their minimum-image displacements as a function of pos.

    static = False: exactly E edges (one host synchronisation to read E).
    static = True : cap_edges rows, no host synchronisation (CUDA-graph capturable).  Rows beyond E are padding: a
    self edge of the last atom with the CONSTANT displacement (cutoff, 0, 0).  There x = 1, and the envelope
    (1-x)^3 p(x) vanishes together with its first and second derivative, so value, gradient and double backward of a
    padded row are exactly zero (the edge MLPs have no bias) and no gradient reaches pos through it."""
    if not static:
        nl.check()
        ei = nl.edge_index()
        dst, src = ei[0], ei[1]
        ep = nl.edge_pair[:nl.n_edges].long()
        valid = None
    else:
        E = nl.cap_edges
        ei = torch.zeros(2, E, dtype=torch.int64, device=pos.device)
        L.check(L.load().nn_nbr_edge_index(C.byref(nl.struct), ei.data_ptr(), E, _stream()), 'nn_nbr_edge_index')
        # a list that outgrew its capacity was not (re)written: expose NO valid rows (the step then computes exact zeros from
        # padding instead of indexing with stale / uninitialised entries) - the caller reads the overflow flag afterwards
        n_valid = torch.where(nl.status[L.ST_EDGE_OVERFLOW] != 0, torch.zeros_like(nl.status[L.ST_N_EDGES]), nl.status[L.ST_N_EDGES])
        valid = torch.arange(E, device=pos.device) < n_valid
        dst = torch.where(valid, ei[0], N - 1)
        src = torch.where(valid, ei[1], N - 1)
        ep = torch.where(valid, nl.edge_pair[:E].long(), 0)
        ei = torch.stack([dst, src])
    sign = torch.where(ep < 0, -1.0, 1.0).to(torch.float32).unsqueeze(1)
    disp0 = nl.pair_disp[(ep & 0x7fffffff)] * sign
    raw = pos.detach()[dst] - pos.detach()[src]
    disp = pos[dst] - pos[src] - (raw - disp0)            # minimum image with a constant lattice shift
    return ei, dst, src, disp, valid


def differentiable_forward(model, z, pos, cell, batch, static_nl=None):
    """NewtonNet.forward with create_graph semantics (reference models/newtonnet.py:74-104 in train mode).
    static_nl: a NeighborList of fixed capacity already rebuilt for `pos` on the current stream -> the whole forward
    has static shapes and no host synchronisation (GraphedTrainingStep)."""
    from newtonnet_b200.engine import get_engine
    from newtonnet_b200.models.output import CustomOutputSet
    props = list(model.output_properties)
    if 'energy' not in props:
        raise RuntimeError("output_properties must contain 'energy'")
    for key in props:
        if key not in ('energy', 'gradient_force', 'direct_force', 'hessian'):
            raise NotImplementedError(f"the differentiable path supports energy, gradient_force, direct_force and hessian, not '{key}'")
    if not pos.is_cuda:
        raise RuntimeError('newtonnet_b200: inputs must be CUDA tensors - there is no CPU fallback')
    if pos.dtype != torch.float32 or next(model.parameters()).dtype != torch.float32:
        raise NotImplementedError('the training path computes in fp32: cast the model and inputs to float32')
    dev = pos.device
    cutoff = model.cutoff
    N, F = pos.shape[0], L.NN_F
    _image_cache.clear()                                    # operand images live for one step
    if model.embedding_layers.requires_dr and pos.is_leaf and not pos.requires_grad:
        pos.requires_grad = True
    # ---- edges (reference order) from the cell-list kernel; image shifts are constants of the graph
    static = static_nl is not None
    nl = static_nl if static else get_engine(dev).checked_neighbor_list(pos, cell, batch, cutoff)   # regrows on overflow
    ei, dst, src, disp, valid = _edges(nl, pos, N, static)
    if static:
        pad = torch.cat([torch.full((1, 1), float(cutoff), dtype=torch.float32, device=dev),
                         torch.zeros(1, 2, dtype=torch.float32, device=dev)], 1)      # fill kernels: no H2D copy while capturing
        disp = torch.where(valid.unsqueeze(1), disp, pad)
    seg_dst = Segments(dst, N, grouped=True, valid=valid)
    seg_src = Segments(src, N, grouped=False if static else None, valid=valid)
    d = disp.norm(dim=1, keepdim=True)
    u = disp / d
    x = d / cutoff
    emb = model.embedding_layers
    freq = emb.edge_embedding.embedding.frequencies
    rbf = _envelope(x) * torch.sin(freq * x) / x
    rbf_pad = Fn.pad(rbf, (0, F - rbf.shape[1]))
    a = Fn.embedding(z, emb.node_embedding.weight, padding_idx=0)
    f = torch.zeros(N, 3 * F, dtype=torch.float32, device=dev)
    for layer in model.interaction_layers:
        n0, n2 = layer.message_nodepart[0], layer.message_nodepart[2]
        mn = linear(Fn.silu(linear(a, n0.weight, n0.bias)), n2.weight, n2.bias)
        # K = 20 contraction through the same fp32-faithful GEMM (zero-padded to K = 128): a library matmul may
        # silently run in single-pass TF32 (TORCH_ALLOW_TF32_CUBLAS_OVERRIDE), which breaks gradient parity
        me = Gemm.apply(rbf_pad, Fn.pad(layer.message_edgepart.weight.t(), (0, 0, 0, F - rbf.shape[1])))
        m = me * Gather.apply(mn, seg_dst) * Gather.apply(mn, seg_src)
        a = a + SegmentSum.apply(m, seg_dst)
        e1 = linear(Fn.silu(linear(m, layer.equiv_message1[0].weight)), layer.equiv_message1[2].weight)
        e2 = linear(Fn.silu(linear(m, layer.equiv_message2[0].weight)), layer.equiv_message2[2].weight)
        fj = Gather.apply(f, seg_src).view(-1, 3, F)
        vec = e1.unsqueeze(1) * u.unsqueeze(2) + e2.unsqueeze(1) * fj
        f = f + SegmentSum.apply(vec.reshape(-1, 3 * F), seg_dst)
        g = linear(f.view(3 * N, F), layer.equiv_update.weight).view(N, 3, F)
        a = a + (f.view(N, 3, F) * g).sum(1)
        if layer.layer_norm is not None:
            a = Fn.layer_norm(a, (F,), layer.layer_norm.weight, layer.layer_norm.bias, layer.layer_norm.eps)
    k = props.index('energy')
    head, scaler = model.output_layers[k].layers, model.scalers[k]
    h = Fn.silu(linear(a, head[0].weight, head[0].bias))
    h = Fn.silu(linear(h, head[2].weight, head[2].bias))
    o = (h * head[4].weight).sum(1, keepdim=True) + head[4].bias           # 128 -> 1: exact fp32 reduction
    e_atom = o * scaler.scale(z) + scaler.shift(z)
    energy = torch.zeros(cell.shape[0], dtype=torch.float32, device=dev).index_add(0, batch, e_atom.reshape(-1))
    out = CustomOutputSet(z=z, pos=pos, cell=cell, batch=batch, edge_index=ei, atom_node=a, force_node=f.view(N, 3, F),
                          displacement=torch.eye(3, device=dev).repeat(cell.shape[0], 1, 1))
    for key in props:
        if key == 'energy':
            out.energy = energy
        elif key == 'direct_force':
            kd = props.index(key)
            dl, ds = model.output_layers[kd].layers, model.scalers[kd]
            hd = Fn.silu(linear(a, dl[0].weight, dl[0].bias))
            hd = linear(Fn.silu(linear(hd, dl[2].weight, dl[2].bias)), dl[4].weight, dl[4].bias)
            out.direct_force = (hd.unsqueeze(1) * f.view(N, 3, F)).sum(-1) * ds.scale(z)
        elif key == 'hessian':
            # reference models/output.py:141-152 (vmap over unit vectors): one reverse pass per row of the 3N x 3N matrix
            if not hasattr(out, 'pos_grad') or not out.pos_grad.requires_grad:
                raise RuntimeError("'hessian' needs 'gradient_force' evaluated before it with create_graph=True "
                                   "(MLAseCalculator sets this, reference utils/ase_interface.py:125-128)")
            flat = out.pos_grad.reshape(-1)
            keep = bool(model.output_layers[props.index(key)].create_graph)
            rows = [torch.autograd.grad(flat[r], pos, retain_graph=True, create_graph=False)[0] for r in range(flat.numel())]
            out.hessian = torch.stack(rows).reshape(N, 3, N, 3)
            if not keep:
                out.hessian = out.hessian.detach()
        else:
            create = bool(model.output_layers[props.index(key)].create_graph) or 'hessian' in props
            out.pos_grad, = torch.autograd.grad(energy, pos, torch.ones_like(energy), create_graph=create,
                                                retain_graph=create)
            out.gradient_force = -out.pos_grad
    return out


# ----------------------------------------------------------------------------- training step (config 5)
def allreduce_gradients(params, group=None, flags=None):
    """Data-parallel gradient averaging: ONE all-reduce of a flat fp32 bucket (401,155 floats = 1.6 MB for
    the default model) instead of one per tensor; missing gradients (the dead layer-0 equiv_message2,
    the frozen frequencies) travel as zeros so every rank reduces the same layout.
    flags: optional 1-D float tensor of per-rank condition flags appended to the bucket; on return it holds their
    SUM over ranks (so every rank takes the same decision on a condition raised by one of them)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return None
    parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params]
    n_flags = 0 if flags is None else flags.numel()
    if n_flags:
        parts.append(flags.reshape(-1).to(parts[0].dtype))
    flat = torch.cat(parts)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat, group=group)
        if n_flags:
            flags.copy_(flat[-n_flags:].reshape(flags.shape))
        flat /= world
    if n_flags:
        flat = flat[:-n_flags]
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat


def training_step(model, optimizer, z, pos, cell, batch, e_target, f_target, force_weight=50.0, clip_grad=1.0,
                  group=None):
    """One step of reference train/trainer.py:303-313: forward (create_graph), loss = MSE(E) + w MSE(F)
    (train/loss.py:48), backward, gradient all-reduce across the data-parallel group, clip, optimizer step."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    pos = pos.detach().clone().requires_grad_(True)
    out = model(z, pos, cell, batch)
    loss = Fn.mse_loss(out.energy, e_target) + force_weight * Fn.mse_loss(out.gradient_force, f_target)
    loss.backward()
    allreduce_gradients(model.parameters(), group)
    if clip_grad and clip_grad > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_grad)
    optimizer.step()
    return loss.detach()


class GraphedTrainingStep:
    """training_step with forward + double backward replayed as ONE CUDA graph (static batch shape).

    Removes every host synchronisation and all Python / autograd dispatch from the step.  Measured on config 5 (100 x 21
    atoms, B200): 18.2 ms against 18.5 ms eager - the step is bound by ~1,400 small kernels on the GPU, not by the host,
    so this is an option (busy hosts, many ranks per host), not the default.  Shapes are made static by
    padding the edge list to the neighbour list's capacity (see `_edges`); the neighbour rebuild, the forward, the loss
    and loss.backward() are captured once, every later call copies the batch into the static buffers and replays.
    Gradient all-reduce, clipping and the optimizer step stay outside the graph (a handful of launches).  Non-periodic
    batches use the exact bound sum n_b (n_b - 1) as edge capacity; periodic ones probe and add headroom, and a replay
    that overflowed raises before the optimizer step (rebuild the object with a larger `cap_edges`)."""

    def __init__(self, model, optimizer, z, pos, cell, batch, e_target, f_target, force_weight=50.0, clip_grad=1.0,
                 group=None, cap_edges=None):
        from newtonnet_b200.engine import NeighborList, get_engine
        dev = pos.device
        self.model, self.optimizer, self.group = model, optimizer, group
        self.force_weight, self.clip_grad = float(force_weight), clip_grad
        self.z, self.batch = z.clone().to(torch.int64), batch.clone().to(torch.int64)
        self.pos = pos.detach().clone().to(torch.float32).contiguous()
        self.cell = cell.detach().clone().to(torch.float32).reshape(-1, 3, 3).contiguous()
        self.e_target, self.f_target = e_target.detach().clone(), f_target.detach().clone()
        self.engine = get_engine(dev)
        self.lib = L.load()
        if cap_edges is None:
            if bool((self.cell == 0).all()):
                n = torch.bincount(self.batch, minlength=self.cell.shape[0])
                cap_edges = int((n * (n - 1)).sum().item())
            else:
                probe = self.engine.neighbor_list(self.pos, self.cell, self.batch, model.cutoff)
                cap_edges = int(probe.check()[L.ST_N_EDGES] * 1.25) + 64
        cap_edges = max(int(cap_edges) + int(cap_edges) % 2, 2)
        self.nl = NeighborList(self.engine, self.pos, self.cell, self.batch, cap_edges=cap_edges)
        model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up on a side stream, as graph capture requires
            for _ in range(2):
                optimizer.zero_grad(set_to_none=True)
                self._forward_backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        self.lib.nn_launch_count(1)
        with torch.cuda.graph(self.graph):
            self.loss = self._forward_backward()
        self.kernels_per_replay = int(self.lib.nn_launch_count(0))      # this library's kernels recorded in the graph
        self.replays = 0

    def _forward_backward(self):
        s = _stream()
        L.check(self.lib.nn_nbr_count(C.byref(self.nl.struct), self.model.cutoff, s), 'nn_nbr_count')
        L.check(self.lib.nn_nbr_fill(C.byref(self.nl.struct), self.model.cutoff, s), 'nn_nbr_fill')
        pos = self.pos.clone().requires_grad_(True)
        out = differentiable_forward(self.model, self.z, pos, self.cell, self.batch, static_nl=self.nl)
        loss = Fn.mse_loss(out.energy, self.e_target) + self.force_weight * Fn.mse_loss(out.gradient_force, self.f_target)
        loss.backward()
        return loss.detach()

    def __call__(self, z, pos, cell, batch, e_target, f_target):
        if pos.shape != self.pos.shape or cell.reshape(-1, 3, 3).shape != self.cell.shape:
            raise ValueError('GraphedTrainingStep was captured for a different batch shape')
        self.z.copy_(z); self.batch.copy_(batch); self.pos.copy_(pos.detach()); self.cell.copy_(cell.detach().reshape(-1, 3, 3))
        self.e_target.copy_(e_target); self.f_target.copy_(f_target)
        self.graph.replay()
        self.replays += 1
        # a batch that outgrew the captured edge capacity leaves stale rows in the replayed graph: the flag travels with
        # the gradient bucket so that every rank sees it, and it is read (one small D2H copy) BEFORE the optimizer moves
        flag = (self.nl.status[L.ST_EDGE_OVERFLOW:L.ST_EDGE_OVERFLOW + 1] != 0).to(torch.float32)
        allreduce_gradients(self.model.parameters(), self.group, flags=flag)
        if float(flag.item()) > 0:
            self.optimizer.zero_grad(set_to_none=True)
            need = int(self.nl.status[L.ST_EDGE_OVERFLOW].item())
            raise RuntimeError(f'edge capacity {self.nl.cap_edges} of the captured training graph overflowed on some rank '
                               f'(this rank needs {need}): the step was NOT applied; rebuild GraphedTrainingStep with a larger cap_edges')
        if self.clip_grad and self.clip_grad > 0:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip_grad)
        self.optimizer.step()
        return self.loss.clone()

    def check(self):
        """Status words of the last replay (one host synchronisation): raises when the edge capacity overflowed."""
        st = self.nl.check()
        if st[L.ST_EDGE_OVERFLOW]:
            raise RuntimeError(f'edge capacity {self.nl.cap_edges} overflowed ({st[L.ST_EDGE_OVERFLOW]} needed)')
        return st
