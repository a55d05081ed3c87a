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
def _gemm_raw(X, B):
    lib = L.load()
    M = X.shape[0]
    Y = torch.empty(M, 128, dtype=torch.float32, device=X.device)
    if M == 0:
        return Y
    img = torch.empty(L.NN_B_IMAGE_FLOATS, dtype=torch.float32, device=X.device)
    L.check(lib.nn_gemm128_prepare_b(B.data_ptr(), img.data_ptr(), _stream()), 'nn_gemm128_prepare_b')
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
        return _gemm_raw(_c(X), _c(B))

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

    def __init__(self, idx, n_rows):
        self.idx = idx.to(torch.int32).contiguous()
        self.n_rows = int(n_rows)
        counts = torch.bincount(idx, minlength=n_rows)
        self.row_ptr = torch.zeros(n_rows + 1, dtype=torch.int32, device=idx.device)
        self.row_ptr[1:] = torch.cumsum(counts, 0)
        if bool((idx[1:] >= idx[:-1]).all()) if idx.numel() > 1 else True:
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


def differentiable_forward(model, z, pos, cell, batch):
    """NewtonNet.forward with create_graph semantics (reference models/newtonnet.py:74-104 in train mode)."""
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
    if model.embedding_layers.requires_dr and pos.is_leaf and not pos.requires_grad:
        pos.requires_grad = True
    # ---- edges (reference order) from the cell-list kernel; image shifts are constants of the graph
    nl = get_engine(dev).neighbor_list(pos, cell, batch, cutoff)
    nl.check()
    ei = nl.edge_index()
    dst, src = ei[0], ei[1]
    ep = nl.edge_pair[:nl.n_edges].long()
    sign = torch.where(ep < 0, -1.0, 1.0).to(torch.float32).unsqueeze(1)
    disp0 = nl.pair_disp[(ep & 0x7fffffff)] * sign
    raw = pos.detach()[dst] - pos.detach()[src]
    disp = pos[dst] - pos[src] - (raw - disp0)            # minimum image with a constant lattice shift
    seg_dst, seg_src = Segments(dst, N), Segments(src, N)
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
def allreduce_gradients(params, group=None):
    """Data-parallel gradient averaging: ONE all-reduce of a flat fp32 bucket (401,155 floats = 1.6 MB for
    the default model) instead of one per tensor; missing gradients (the dead layer-0 equiv_message2,
    the frozen frequencies) travel as zeros so every rank reduces the same layout."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return None
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat /= world
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
