"""Multi-GPU evaluation: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

Two partitionings, as SURVEY.md section 8e lays out (the reference itself is single-process):

* independent molecule batches - `shard_batch` splits the systems of a batch into contiguous,
  atom-balanced ranges; every rank evaluates its range with the ordinary single-GPU path, there is no
  data-path collective (edges never cross systems, reference layers/representations.py:74-77);

* one large periodic box - `DomainDecomposition` splits space into a px x py x pz grid of bricks.  A
  rank owns the atoms inside its brick and keeps ghost copies of every atom within the cutoff of it.
  Pairs with at least one owned endpoint are evaluated locally (owned-ghost pairs on both sides, ~10 %
  redundant edge-MLP work for 50 A bricks), so every per-atom sum of an owned atom is complete locally
  and the only communication is owner -> ghost copies of per-atom feature rows, three times per
  direction: forward mn(l) [+ f(l-1)], reverse dfb(l) + abar(l).  Ghost rows are ordered by owner rank,
  so a peer's rows land contiguously: the receive side needs no unpack kernel.
  Energy, virial and forces are completed with one all-reduce each.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1), 6: (3, 2, 1), 8: (2, 2, 2)}


# ----------------------------------------------------------------------------- molecule batches
def shard_batch(z, pos, cell, batch, rank, world):
    """Contiguous range of systems for `rank`, balanced by atom count (numpy or torch inputs).
    Returns (z, pos, cell, batch_local, system_slice) with batch re-based to start at 0."""
    b = np.asarray(batch.cpu() if torch.is_tensor(batch) else batch)
    n_sys = cell.shape[0]
    counts = np.bincount(b, minlength=n_sys)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = [int(np.searchsorted(csum, total * r / world, side='left')) for r in range(world)] + [n_sys]
    bounds[0] = 0
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    s0, s1 = bounds[rank], bounds[rank + 1]
    a0, a1 = int(csum[s0]), int(csum[s1])
    return z[a0:a1], pos[a0:a1], cell[s0:s1], batch[a0:a1] - s0, slice(s0, s1)


# ----------------------------------------------------------------------------- spatial decomposition
class HaloPlan:
    """Host-side description of one rank's brick: who is owned, who is a ghost, what to send to whom."""

    def __init__(self, pos, cell, rank, world, cutoff, grid=None, skin=0.0):
        pos = np.asarray(pos, dtype=np.float64)
        cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        if np.count_nonzero(cell - np.diag(np.diag(cell))) or np.any(np.diag(cell) <= 0):
            raise ValueError('domain decomposition needs an orthorhombic periodic cell')
        grid = tuple(grid) if grid is not None else GRIDS.get(world)
        if grid is None or int(np.prod(grid)) != world:
            raise ValueError(f'no brick grid for world size {world}')
        Ld = np.diag(cell)
        width = Ld / np.array(grid)
        halo = cutoff * 1.0001 + skin + 16 * np.finfo(np.float32).eps * np.abs(pos).max(initial=0.0)
        for d in range(3):
            if grid[d] > 1 and width[d] < halo:
                raise ValueError('bricks must be at least one cutoff wide')
        self.grid, self.rank, self.world = grid, rank, world
        frac = pos / Ld
        frac -= np.floor(frac)
        wrapped = frac * Ld
        ix = np.minimum((frac * np.array(grid)).astype(np.int64), np.array(grid) - 1)
        owner = (ix[:, 0] * grid[1] + ix[:, 1]) * grid[2] + ix[:, 2]
        self.owner = owner
        all_idx = np.arange(len(pos))

        def ghosts_of(r):
            """Atoms not owned by r within `halo` of r's brick (per-axis periodic distance)."""
            rc = np.array([r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2]])
            lo, hi = rc * width, (rc + 1) * width
            near = np.ones(len(pos), dtype=bool)
            for d in range(3):
                if grid[d] == 1:
                    continue
                x = wrapped[:, d]
                below = np.minimum(np.abs(lo[d] - x), Ld[d] - np.abs(lo[d] - x))
                above = np.minimum(np.abs(x - hi[d]), Ld[d] - np.abs(x - hi[d]))
                inside = (x >= lo[d]) & (x < hi[d])
                near &= inside | (np.minimum(below, above) <= halo)
            g = all_idx[near & (owner != r)]
            return g[np.lexsort((g, owner[g]))]            # grouped by owner rank, ascending global index

        self.owned = all_idx[owner == rank]
        self.ghost = ghosts_of(rank)
        self.n_owned, self.n_ghost = len(self.owned), len(self.ghost)
        self.local_to_global = np.concatenate([self.owned, self.ghost])
        g_owner = owner[self.ghost]
        self.recv_counts = [int((g_owner == r).sum()) for r in range(world)]
        # rows this rank sends: its owned atoms that are ghosts of peer s, in s's ghost order
        local_of = -np.ones(len(pos), dtype=np.int64)
        local_of[self.owned] = np.arange(self.n_owned)
        send = []
        self.send_counts = []
        self.send_row_offset = []        # first row of this rank's block in peer s's ghost order
        for s in range(world):
            if s == rank:
                self.send_counts.append(0)
                self.send_row_offset.append(0)
                continue
            gs = ghosts_of(s)
            mine = gs[owner[gs] == rank]
            send.append(local_of[mine])
            self.send_counts.append(len(mine))
            self.send_row_offset.append(int((owner[gs] < rank).sum()))
        self.send_index = np.concatenate(send).astype(np.int32) if send else np.zeros(0, np.int32)


class HaloExchange:
    """Owner -> ghost copy of per-atom rows: pack kernel + all_to_all_single into the ghost tail."""

    def __init__(self, plan, device, group=None):
        self.plan, self.group = plan, group
        self.device = torch.device(device)
        self.send_index = torch.from_numpy(plan.send_index).to(self.device)
        self.n_send = int(plan.send_index.shape[0])
        self._buf = {}

    def _pack(self, rows, width):
        key = width
        buf = self._buf.get(key)
        if buf is None or buf.shape[0] < self.n_send:
            buf = torch.empty(max(self.n_send, 1), width, dtype=torch.float32, device=self.device)
            self._buf[key] = buf
        out = buf[:self.n_send]
        if self.n_send == 0:
            return out
        if rows.is_cuda:
            L.check(L.load().nn_halo_pack(rows.data_ptr(), self.send_index.data_ptr(), self.n_send, width,
                                          out.data_ptr(), torch.cuda.current_stream().cuda_stream), 'nn_halo_pack')
        else:   # host tensors: used only by the gloo tests of the plan / exchange logic
            torch.index_select(rows, 0, self.send_index.long(), out=out)
        return out

    def exchange(self, rows):
        """rows: [n_local, width] float32 (contiguous); ghost rows [n_owned:] are overwritten."""
        p = self.plan
        if p.world == 1:
            return
        width = rows.shape[1]
        send = self._pack(rows, width)
        recv = rows[p.n_owned:]
        if dist.get_backend(self.group) == 'nccl':
            dist.all_to_all_single(recv, send, output_split_sizes=p.recv_counts, input_split_sizes=p.send_counts,
                                   group=self.group)
        else:
            ops, so, ro = [], 0, 0
            for r in range(p.world):
                if p.send_counts[r]:
                    ops.append(dist.P2POp(dist.isend, send[so:so + p.send_counts[r]], r, group=self.group))
                if p.recv_counts[r]:
                    ops.append(dist.P2POp(dist.irecv, recv[ro:ro + p.recv_counts[r]], r, group=self.group))
                so += p.send_counts[r]; ro += p.recv_counts[r]
            for req in (dist.batch_isend_irecv(ops) if ops else []):
                req.wait()


class PeerHaloExchange:
    """Owner -> ghost copy over NVLink peer memory (csrc/p2p.cu): the pack kernel stores rows directly into
    the peers' landing buffers and raises per-source flags; the receiver spins on its flags and copies
    its landing buffer into the ghost tail.  No NCCL call on the data path.  Collective: every rank of
    the group must construct it (and call `exchange`) in the same order."""

    MAX_WIDTH = 3 * L.NN_F

    def __init__(self, plan, device, group=None):
        self.plan, self.group, self.device = plan, group, torch.device(device)
        self.lib = L.load()
        self.rank, self.world = plan.rank, plan.world
        self.epoch = 0
        self.send_index = torch.from_numpy(plan.send_index).to(self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.done = torch.zeros(1, **i32)
        self.status = torch.zeros(1, **i32)
        self.expect = torch.tensor([1 if c > 0 else 0 for c in plan.recv_counts], **i32)
        nbytes = max(plan.n_ghost, 1) * self.MAX_WIDTH * 4
        self.local = []
        for _ in range(3):           # two landing buffers (epoch parity) and the flag array
            p = C.c_void_p()
            L.check(self.lib.nn_p2p_alloc(nbytes if len(self.local) < 2 else 4 * max(self.world, 1), C.byref(p)), 'nn_p2p_alloc')
            self.local.append(p.value)
        handles = []
        for p in self.local:
            h = C.create_string_buffer(64)
            L.check(self.lib.nn_p2p_get_handle(p, h), 'nn_p2p_get_handle')
            handles.append(h.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self.peers = [s for s in range(self.world) if s != self.rank and plan.send_counts[s] > 0]
        self.opened = {}
        for s in range(self.world):
            if s == self.rank or (plan.send_counts[s] == 0):
                continue
            ptrs = []
            for raw in gathered[s]:
                q = C.c_void_p()
                L.check(self.lib.nn_p2p_open_handle(C.create_string_buffer(raw, 64), C.byref(q)), 'nn_p2p_open_handle')
                ptrs.append(q.value)
            self.opened[s] = ptrs
        n = len(self.peers)
        begins, ends, off = [], [], 0
        counts_nonself = [(s, plan.send_counts[s]) for s in range(self.world) if s != self.rank]
        pos = {}
        for s, c in counts_nonself:      # send_index is ordered by destination rank
            pos[s] = (off, off + c)
            off += c
        self._arrays = []
        for par in range(2):
            landing = (C.c_void_p * max(n, 1))(*[self.opened[s][par] for s in self.peers])
            flags = (C.c_void_p * max(n, 1))(*[self.opened[s][2] for s in self.peers])
            self._arrays.append((landing, flags))
        self._row_offset = (C.c_int32 * max(n, 1))(*[plan.send_row_offset[s] for s in self.peers])
        self._begin = (C.c_int32 * max(n, 1))(*[pos[s][0] for s in self.peers])
        self._end = (C.c_int32 * max(n, 1))(*[pos[s][1] for s in self.peers])
        if self.world > 1:
            dist.barrier(group=group)

    def exchange(self, rows):
        p = self.plan
        if p.world == 1:
            return
        self.epoch += 1
        par = self.epoch & 1
        width = rows.shape[1]
        s = torch.cuda.current_stream().cuda_stream
        if self.peers:
            landing, flags = self._arrays[par]
            L.check(self.lib.nn_halo_push(rows.data_ptr(), self.send_index.data_ptr(), width, len(self.peers), landing, flags,
                                          self._row_offset, self._begin, self._end, self.rank, self.epoch,
                                          self.done.data_ptr(), s), 'nn_halo_push')
        if p.n_ghost:
            L.check(self.lib.nn_halo_wait(self.local[2], self.expect.data_ptr(), self.world, self.epoch,
                                          self.status.data_ptr(), s), 'nn_halo_wait')
            L.check(self.lib.nn_copy_d2d(rows[p.n_owned:].data_ptr(), self.local[par], p.n_ghost * width * 4, s), 'nn_copy_d2d')

    def check(self):
        st = int(self.status.item())
        if st:
            raise RuntimeError(f'halo exchange timed out waiting for rank {st - 1}')

    def close(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        for ptrs in self.opened.values():
            for q in ptrs:
                self.lib.nn_p2p_close_handle(q)
        for q in self.local:
            self.lib.nn_p2p_free(q)
        self.opened, self.local = {}, []


class DomainDecomposition:
    """Energy / forces / stress of ONE periodic box across the ranks of a process group.

    dd = DomainDecomposition(model); out = dd(z, pos, cell)   (same full inputs on every rank)
    -> CustomOutputSet with energy [1], gradient_force [N,3] (complete on every rank), stress, virial.
    """

    def __init__(self, model, group=None, grid=None, skin=1.0, transport='p2p'):
        """skin (A): the brick/ghost plan is kept while no atom has moved more than skin/2 since it was
        made (the ghost shell is cutoff + skin thick); the neighbour list itself is rebuilt every call.
        transport: 'p2p' = pack kernel storing into peer memory over NVLink (csrc/p2p.cu), 'nccl' =
        pack kernel + all_to_all_single."""
        self.model, self.group, self.grid, self.skin = model, group, grid, float(skin)
        self.transport = transport
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.plan = None
        self._ws = None
        self._nl = None
        self._state = None
        self.n_plans = 0

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.05) + 256, dtype=torch.uint8, device=device)
        return self._ws

    def _view(self, ws, ptr, rows, width):
        off = ptr - ws.data_ptr()
        return ws[off:off + rows * width * 4].view(torch.float32).view(rows, width)

    def _replan(self, z, pos, cell3, cutoff, dev):
        plan = HaloPlan(pos.detach().cpu().numpy(), cell3[0].detach().cpu().numpy(), self.rank, self.world,
                        cutoff, self.grid, skin=self.skin)
        self.plan = plan
        self.n_plans += 1
        l2g = torch.from_numpy(plan.local_to_global).to(dev)
        if self._state is not None and hasattr(self._state['halo'], 'close'):
            self._state['halo'].close()
        use_p2p = self.transport == 'p2p' and self.world > 1 and dist.get_backend(self.group) == 'nccl'
        halo = PeerHaloExchange(plan, dev, self.group) if use_p2p else HaloExchange(plan, dev, self.group)
        self._state = dict(l2g=l2g, halo=halo, pos_ref=pos.detach().clone(),
                           z_l=z.to(torch.int64)[l2g].contiguous(), n=pos.shape[0],
                           batch_l=torch.zeros(len(plan.local_to_global), dtype=torch.int64, device=dev))
        self._nl = None

    def __call__(self, z, pos, cell, want_virial=True):
        from newtonnet_b200.engine import NeighborList, _stream, get_engine
        from newtonnet_b200.models.output import CustomOutputSet
        model = self.model
        dev = pos.device
        engine = get_engine(dev)
        lib = engine.lib
        pack = model._weight_pack(dev)
        N = pos.shape[0]
        cell3 = cell.reshape(-1, 3, 3)
        if cell3.shape[0] != 1:
            raise ValueError('DomainDecomposition evaluates one periodic system')
        st = self._state
        stale = st is None or st['n'] != N
        if not stale:   # one scalar read-back decides whether the cached plan still covers the cutoff
            moved = float((pos.detach() - st['pos_ref']).square().sum(1).max().sqrt())
            flag = torch.tensor([1.0 if moved > 0.5 * self.skin else 0.0], device=dev)
            if self.world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            stale = bool(flag.item() > 0)
        if stale:
            self._replan(z, pos, cell3, pack.cutoff, dev)
            st = self._state
        plan, halo, l2g, z_l, batch_l = self.plan, st['halo'], st['l2g'], st['z_l'], st['batch_l']
        pos_l = pos.detach().to(torch.float32)[l2g].contiguous()
        cell_l = cell3.detach().to(torch.float32).contiguous()
        n_local, n_owned = int(l2g.shape[0]), plan.n_owned
        s = _stream()

        # ---- neighbour list over owned + ghost atoms (ghost-ghost pairs dropped); capacities are kept
        def build(cap):
            nl = NeighborList(engine, pos_l, cell_l, batch_l, cap_edges=cap)
            nl.struct.n_owned = n_owned
            return nl

        def run_nbr(nl, fill=True):
            L.check(lib.nn_nbr_count(C.byref(nl.struct), pack.cutoff, s), 'nn_nbr_count')
            if fill:
                L.check(lib.nn_nbr_fill(C.byref(nl.struct), pack.cutoff, s), 'nn_nbr_fill')

        if self._nl is None:
            probe = build(0)
            run_nbr(probe, fill=False)
            cap = int(probe.check()[L.ST_N_EDGES] * 1.08) + 64
            self._nl = build(cap + cap % 2)
        nl = self._nl
        nl.rebind(pos_l, cell_l, batch_l)
        nl.n_edges = None
        run_nbr(nl)
        n_layers = pack.n_layers

        # ---- phased evaluation with halo exchanges
        f32 = dict(dtype=torch.float32, device=dev)
        energy = torch.zeros(1, **f32); forces_l = torch.zeros(max(n_owned, 1), 3, **f32)
        virial = torch.zeros(1, 3, 3, **f32); stress = torch.zeros(1, 3, 3, **f32)
        nbytes = lib.nn_eval_workspace_bytes(n_local, 1, nl.cap_pairs, n_layers, 1)
        ws = self._workspace(nbytes, dev)
        a = L.EvalArgs()
        a.nbr, a.w, a.z = C.pointer(nl.struct), C.pointer(pack.struct), z_l.data_ptr()
        a.want_forces, a.want_virial, a.n_owned = 1, int(want_virial), n_owned
        a.energy, a.forces = energy.data_ptr(), forces_l.data_ptr()
        a.virial, a.stress = (virial.data_ptr(), stress.data_ptr()) if want_virial else (None, None)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()

        def phase(ph, l=0):
            L.check(lib.nn_eval_phase(C.byref(a), ph, l, s), f'nn_eval_phase({ph},{l})')

        def buf(which, l, width):
            return self._view(ws, lib.nn_eval_buffer(C.byref(a), which, l), n_local, width)

        phase(L.PH_BEGIN)
        for l in range(n_layers):
            phase(L.PH_FWD_NODE, l)
            halo.exchange(buf(L.BUF_MN, l, L.NN_F))
            if l > 0:
                halo.exchange(buf(L.BUF_F_OUT, l - 1, 3 * L.NN_F))
            phase(L.PH_FWD_PAIR, l)
        phase(L.PH_HEAD)
        phase(L.PH_BWD_SEED)
        for l in reversed(range(n_layers)):
            phase(L.PH_BWD_NODE, l)
            halo.exchange(buf(L.BUF_DFB, 0, 3 * L.NN_F))
            halo.exchange(buf(L.BUF_ABAR, 0, L.NN_F))
            phase(L.PH_BWD_PAIR, l)
        phase(L.PH_FINISH)

        # ---- complete the sums across ranks: ONE all-reduce of [forces | energy, virial, stress (hi, lo) | flags]
        status_dev = nl.status.to(torch.float32)
        small = torch.cat([energy.double(), virial.double().reshape(-1), stress.double().reshape(-1)])
        hi = small.float()
        lo = (small - hi.double()).float()                       # fp64 partial sums travel as two fp32 parts
        over_flag = (status_dev[L.ST_EDGE_OVERFLOW:L.ST_EDGE_OVERFLOW + 1] != 0).float()
        buf = torch.zeros(3 * N + 2 * 19 + 1, **f32)
        buf[:3 * N].view(N, 3)[l2g[:n_owned]] = forces_l[:n_owned]
        buf[3 * N:3 * N + 19] = hi
        buf[3 * N + 19:3 * N + 38] = lo
        buf[3 * N + 38:] = over_flag
        if self.world > 1:
            dist.all_reduce(buf, group=self.group)
        forces = buf[:3 * N].view(N, 3)
        tail = buf[3 * N:].cpu()                                  # the one host synchronisation of the step
        red = tail[:19].double() + tail[19:38].double()
        status = nl.check()
        if hasattr(halo, 'check'):
            halo.check()
        if float(tail[38]) > 0:      # some rank's list outgrew its capacity: resize everywhere and repeat
            self._nl = None
            return self.__call__(z, pos, cell, want_virial)
        dt = pos.dtype
        out = CustomOutputSet(z=z, pos=pos, cell=cell, batch=torch.zeros(N, dtype=torch.int64, device=dev))
        red = red.to(dev)
        out.energy = red[:1].to(dt)
        out.gradient_force = forces.to(dt)
        out.virial = red[1:10].reshape(1, 3, 3).to(dt)
        out.stress = red[10:19].reshape(1, 3, 3).to(dt)
        out.n_owned, out.n_ghost, out.n_local_edges = n_owned, plan.n_ghost, status[L.ST_N_EDGES]
        out._keep = (nl, z_l, pos_l, cell_l, batch_l, halo)
        return out
