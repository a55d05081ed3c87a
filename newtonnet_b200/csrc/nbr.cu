// Neighbour-list build: per-system cell list -> destination-sorted CSR + undirected pair list.
//
// Replaces RadiusGraph.forward (reference newtonnet/layers/representations.py:57-100), which builds a
// dense O(N^2) ordered-pair mesh per system and filters it.  Here candidates come from a grid of cells
// no smaller than the cutoff; every candidate (i, j) is then tested with the reference's own fp32
// arithmetic (nn_min_image / nn_norm3 in common.cuh), so the surviving edge set is identical.  All
// work is HBM/latency-bound integer + fp32 work: warp per atom, coalesced candidate reads, no atomics
// on the output (two-pass count / fill), rows sorted by source index so the order equals the
// reference's (i-major, j ascending).
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

struct NbrWs {
    SysMeta* meta;        // [B]
    unsigned int* bounds; // [B,6] min xyz, max xyz (order-preserving integer encoding, see k_sys_bounds)
    int* atom_cell;       // [N]
    int* cell_count;      // [cap_cells+1]
    int* cell_start;      // [cap_cells+1]
    int* cell_fill;       // [cap_cells]
    int* sorted_atoms;    // [N]
    int* deg;             // [N+1]
    int* fwd_cnt;         // [N+1]
    void* scan_tmp; size_t scan_tmp_bytes;
    size_t total;
};

// ---------------------------------------------------------------- exclusive prefix sum (int32)
// Tiles of 4096 elements: one 1024-thread block scans a tile (4 elements per thread, warp shuffles, one shared-memory hop)
// and emits the tile total; totals are scanned by the same kernel one level up and added back.  One launch for
// n <= 4096 (every list of a small system), three for up to 16.7 M elements.
constexpr int kScanTile = 4096;
__global__ void __launch_bounds__(1024) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int n,
                                                     int* __restrict__ tile_total) {
    __shared__ int s_warp[32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < n ? in[base + k] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        s_warp[lane] = wi - w;                  // exclusive prefix of the warp totals
        if (lane == 31 && tile_total) tile_total[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = s_warp[wid] + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ tile_prefix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_prefix[i / kScanTile];
}
size_t scan_temp_bytes(int n) {
    size_t ints = 0;
    for (int m = nn_ceil_div(n, kScanTile); m > 1; m = nn_ceil_div(m, kScanTile)) ints += (size_t)m;
    return (ints + 1) * sizeof(int);
}
// out may alias in.  Returns the number of kernels launched.
int exclusive_scan(const int* in, int* out, int n, int* tmp, cudaStream_t s) {
    if (n <= 0) return 0;
    const int tiles = nn_ceil_div(n, kScanTile);
    if (tiles == 1) { k_scan_tiles<<<1, 1024, 0, s>>>(in, out, n, nullptr); return 1; }
    k_scan_tiles<<<tiles, 1024, 0, s>>>(in, out, n, tmp);
    int launched = 1 + exclusive_scan(tmp, tmp, tiles, tmp + tiles, s);
    k_scan_add<<<nn_ceil_div(n, 256), 256, 0, s>>>(out, n, tmp);
    return launched + 1;
}

NbrWs carve(void* base, size_t cap, int N, int B) {
    WsCarver c(base, cap);
    NbrWs w;
    int cap_cells = 2 * N + B + 1;
    w.meta = c.take<SysMeta>(B);
    w.bounds = c.take<unsigned int>((size_t)B * 6);
    w.atom_cell = c.take<int>(N);
    w.cell_count = c.take<int>(cap_cells + 1);
    w.cell_start = c.take<int>(cap_cells + 1);
    w.cell_fill = c.take<int>(cap_cells);
    w.sorted_atoms = c.take<int>(N);
    w.deg = c.take<int>(N + 1);
    w.fwd_cnt = c.take<int>(N + 1);
    w.scan_tmp_bytes = scan_temp_bytes(cap_cells + 1 > N + 1 ? cap_cells + 1 : N + 1);
    w.scan_tmp = c.take<char>(w.scan_tmp_bytes);
    w.total = c.off;
    return w;
}

// ---------------------------------------------------------------- system table
__global__ void k_sys_ptr(const int64_t* __restrict__ batch, int N, int B, int* __restrict__ sys_ptr,
                          int* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    long long prev = (i == 0) ? -1 : batch[i - 1];
    long long cur = (i == N) ? B : batch[i];
    if (cur < prev || cur > B || (i < N && cur >= B) || prev < -1) { atomicExch(&status[NN_ST_BATCH_UNSORTED], 1); return; }
    for (long long s = prev + 1; s <= cur; ++s) sys_ptr[s] = i;
}

// Bounding box of every system.  min / max do not depend on the order of the operands, so atomics on an order-preserving
// integer encoding of the floats are deterministic; a warp whose 32 atoms belong to one system (the common case)
// reduces with shuffles and issues six atomics.  (One block per system took 361 us on the 98,304-atom box.)
__device__ __forceinline__ unsigned int f2ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void k_bounds_init(unsigned int* __restrict__ bounds, int B) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * 6) bounds[t] = (t % 6) < 3 ? f2ord(3.0e38f) : f2ord(-3.0e38f);
}
__global__ void k_sys_bounds(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N, int B,
                             unsigned int* __restrict__ bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < N;
    long long b = ok ? batch[i] : -1;
    if (b < 0 || b >= B) b = -1;                        // invalid batch entries are reported by k_sys_ptr
    float v[3] = {0.f, 0.f, 0.f};
    if (ok) { v[0] = pos[3 * i]; v[1] = pos[3 * i + 1]; v[2] = pos[3 * i + 2]; }
    const long long b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0) && b0 >= 0;
    if (uniform) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float lo = v[d], hi = v[d];
            for (int o = 16; o > 0; o >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[6 * b0 + d], f2ord(lo)); atomicMax(&bounds[6 * b0 + 3 + d], f2ord(hi)); }
        }
    } else if (b >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { atomicMin(&bounds[6 * b + d], f2ord(v[d])); atomicMax(&bounds[6 * b + 3 + d], f2ord(v[d])); }
    }
}

// One block: periodic flag over the whole batch (reference: `not (cell == 0).all()`,
// representations.py:86 - if ANY entry of ANY cell is non-zero every system takes the solve path),
// then per-system mode / grid, then an exclusive scan of the grid sizes.
__global__ void __launch_bounds__(1024) k_sys_plan(const float* __restrict__ cell, const int* __restrict__ sys_ptr,
                           const unsigned int* __restrict__ bounds, int B, float cutoff, SysMeta* __restrict__ meta,
                           int* __restrict__ status) {
    __shared__ int s_any;
    __shared__ int s_carry;
    __shared__ int s_scan[1024];
    if (threadIdx.x == 0) { s_any = 0; s_carry = 0; }
    __syncthreads();
    int any = 0;
    for (int k = threadIdx.x; k < B * 9; k += blockDim.x) any |= (cell[k] != 0.0f);
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    const bool periodic = s_any != 0;
    for (int base = 0; base < B; base += blockDim.x) {
        int b = base + threadIdx.x;
        int ncells = 0;
        SysMeta m;
        if (b < B) {
            m.first = sys_ptr[b]; m.count = sys_ptr[b + 1] - sys_ptr[b]; m.nimg = 0;
            const float* h = cell + 9 * b;
            float lo[3] = {ord2f(bounds[6 * b]), ord2f(bounds[6 * b + 1]), ord2f(bounds[6 * b + 2])};
            float hi[3] = {ord2f(bounds[6 * b + 3]), ord2f(bounds[6 * b + 4]), ord2f(bounds[6 * b + 5])};
            if (m.count == 0) { lo[0] = lo[1] = lo[2] = 0.f; hi[0] = hi[1] = hi[2] = 0.f; }
            float amax = 0.f;
            for (int d = 0; d < 3; ++d) amax = fmaxf(amax, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
            // smallest admissible grid cell: cutoff + safety for fp32 rounding of pos_i - pos_j
            double wmin = (double)cutoff * 1.0001 + 16.0 * 1.1920929e-7 * (double)amax;
            for (int k = 0; k < 9; ++k) { m.H[k] = h[k]; m.Hinv[k] = 0.f; }
            for (int d = 0; d < 3; ++d) { m.lo[d] = 0.f; m.wsc[d] = 0.f; m.L[d] = 0.f; m.nc[d] = 1; }
            double ext[3];
            if (!periodic) {
                m.mode = 0;
                for (int d = 0; d < 3; ++d) { ext[d] = fmax((double)hi[d] - (double)lo[d], 1e-6); m.lo[d] = lo[d]; }
            } else {
                bool diag = h[1] == 0.f && h[2] == 0.f && h[3] == 0.f && h[5] == 0.f && h[6] == 0.f && h[7] == 0.f;
                double H[9]; for (int k = 0; k < 9; ++k) H[k] = h[k];
                double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
                if (det == 0.0) atomicExch(&status[NN_ST_SINGULAR_CELL], 1);
                if (diag && h[0] > 0.f && h[4] > 0.f && h[8] > 0.f) {
                    m.mode = 1;
                    for (int d = 0; d < 3; ++d) { m.L[d] = h[4 * d]; ext[d] = h[4 * d]; }
                } else {
                    m.mode = 2;
                    // inverse of cell^T = (cofactor matrix of cell) / det
                    double inv = det != 0.0 ? 1.0 / det : 0.0;
                    double C[9];
                    C[0] = (H[4] * H[8] - H[5] * H[7]); C[1] = -(H[3] * H[8] - H[5] * H[6]); C[2] = (H[3] * H[7] - H[4] * H[6]);
                    C[3] = -(H[1] * H[8] - H[2] * H[7]); C[4] = (H[0] * H[8] - H[2] * H[6]); C[5] = -(H[0] * H[7] - H[1] * H[6]);
                    C[6] = (H[1] * H[5] - H[2] * H[4]); C[7] = -(H[0] * H[5] - H[2] * H[3]); C[8] = (H[0] * H[4] - H[1] * H[3]);
                    for (int k = 0; k < 9; ++k) m.Hinv[k] = (float)(C[k] * inv);
                    // The reference keeps ONE image per ordered pair, n = round((cell^T)^-1 d), and tests |d - cell n| < r_c
                    // (its `cell @ n` quirk), i.e. atom j passes iff it lies within r_c of the point p_i - cell n.  So the
                    // candidates of atom i are the atoms near the query points p_i - cell n for every image n that can occur,
                    // found in a NON-periodic Cartesian grid over the atoms' bounding box; a hit counts only if the reference
                    // formula returns this very n (each pair has exactly one).  |n_a| <= sum_k |Hinv[a][k]| * extent_k.
                    double nmax = 0.0, hsum = 0.0;
                    for (int a = 0; a < 3; ++a) {
                        double bnd = 0.0;
                        for (int k = 0; k < 3; ++k) bnd += fabs(C[3 * a + k] * inv) * fmax((double)hi[k] - (double)lo[k], 0.0);
                        nmax = fmax(nmax, bnd);
                    }
                    for (int k = 0; k < 9; ++k) hsum += fabs(H[k]);
                    if (det != 0.0 && nmax < 2.49) {
                        m.nimg = (int)floor(nmax + 0.501);        // margin: fp32 rounding of (cell^T)^-1 d next to a half-integer
                        wmin += 64.0 * 1.1920929e-7 * (hsum * (m.nimg + 1) + (double)amax);    // fp32 rounding of p_i - cell n
                        for (int d = 0; d < 3; ++d) { ext[d] = fmax((double)hi[d] - (double)lo[d], 1e-6); m.lo[d] = lo[d]; }
                    } else {
                        m.nimg = -1;                  // unwrapped input far outside the cell: single grid cell, all pairs
                        ext[0] = ext[1] = ext[2] = 0.0;
                    }
                }
            }
            long long cap = 2LL * m.count; if (cap < 1) cap = 1;
            int nc[3];
            for (int d = 0; d < 3; ++d) {
                double q = floor(ext[d] / wmin);
                nc[d] = q < 1.0 ? 1 : (q > 1024.0 ? 1024 : (int)q);
            }
            // sparse boxes: coarsen (cells may only get larger) until the grid has <= 2 n_atoms cells
            while ((long long)nc[0] * nc[1] * nc[2] > cap) {
                int dmax = 0;
                if (nc[1] > nc[dmax]) dmax = 1;
                if (nc[2] > nc[dmax]) dmax = 2;
                nc[dmax] -= 1;
            }
            for (int d = 0; d < 3; ++d) {
                m.nc[d] = nc[d];
                m.wsc[d] = (m.mode == 0 || (m.mode == 2 && m.nimg >= 0)) ? (float)((double)nc[d] / ext[d]) : (float)nc[d];
            }
            ncells = nc[0] * nc[1] * nc[2];
        }
        // block exclusive scan of ncells
        s_scan[threadIdx.x] = ncells;
        __syncthreads();
        for (int o = 1; o < blockDim.x; o <<= 1) {
            int v = (threadIdx.x >= o) ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (b < B) { m.cell_off = s_carry + s_scan[threadIdx.x] - ncells; meta[b] = m; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) status[NN_ST_N_CELLS] = s_carry;
}

__device__ __forceinline__ void cell_coords(const SysMeta& m, float x, float y, float z, int c[3]) {
    float p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int v = 0;
        if (m.mode == 0 || (m.mode == 2 && m.nimg >= 0)) {
            v = (int)(((double)p[d] - (double)m.lo[d]) * (double)m.wsc[d]);
        } else if (m.mode == 1) {
            double f = (double)p[d] / (double)m.L[d];
            f -= floor(f);
            v = (int)(f * (double)m.nc[d]);
        }
        c[d] = v < 0 ? 0 : (v >= m.nc[d] ? m.nc[d] - 1 : v);
    }
}

__global__ void k_bin_count(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N,
                            const SysMeta* __restrict__ meta, int* __restrict__ atom_cell,
                            int* __restrict__ cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const SysMeta& m = meta[batch[i]];
    int c[3];
    cell_coords(m, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], c);
    int cid = m.cell_off + (c[0] * m.nc[1] + c[1]) * m.nc[2] + c[2];
    atom_cell[i] = cid;
    atomicAdd(&cell_count[cid], 1);
}

__global__ void k_bin_fill(int N, const int* __restrict__ atom_cell, const int* __restrict__ cell_start,
                           int* __restrict__ cell_fill, int* __restrict__ sorted_atoms) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int cid = atom_cell[i];
    int slot = atomicAdd(&cell_fill[cid], 1);
    sorted_atoms[cell_start[cid] + slot] = i;   // order inside a cell is irrelevant: rows are sorted later
}

// Candidates of one grid cell: every lane tests one atom j with the reference's arithmetic.  IMG >= 0: general cell, the
// hit counts only when the reference's own image vector equals the enumerated one (img_x/y/z).
template <bool FILL, bool IMG>
__device__ __forceinline__ void scan_cell(int i, int lane, const float3 pi, const float* __restrict__ pos, const SysMeta& m,
                                          int s0, int s1, const int* __restrict__ sorted_atoms, float cutoff, int n_owned,
                                          float img_x, float img_y, float img_z, int* __restrict__ row_buf, int& found) {
    for (int s = s0 + lane; s - lane < s1; s += 32) {
        bool pass = false;
        int j = -1;
        if (s < s1) {
            j = sorted_atoms[s];
            if (j != i && (i < n_owned || j < n_owned)) {   // ghost-ghost pairs belong to other ranks
                float3 d = make_float3(__fsub_rn(pi.x, pos[3 * j]), __fsub_rn(pi.y, pos[3 * j + 1]),
                                       __fsub_rn(pi.z, pos[3 * j + 2]));
                float3 n3;
                d = nn_min_image(d, m, &n3);
                pass = nn_norm3(d) < cutoff;
                if (IMG) pass = pass && n3.x == img_x && n3.y == img_y && n3.z == img_z;
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (FILL && pass) {
            int slot = found + __popc(mask & ((1u << lane) - 1u));
            if (slot < NN_MAX_DEGREE) row_buf[slot] = j;
        }
        found += __popc(mask);
    }
}

// Visit every candidate neighbour of atom i once; FILL = false counts, FILL = true collects into smem.
template <bool FILL>
__device__ __forceinline__ int visit_neighbours(int i, int lane, const float* __restrict__ pos,
                                                const SysMeta& m, const int* __restrict__ cell_start,
                                                const int* __restrict__ sorted_atoms, float cutoff,
                                                int n_owned, int* __restrict__ row_buf) {
    const float3 pi = make_float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    int found = 0;
    if (m.mode == 2 && m.nimg >= 0) {
        // general cell: query points q = p_i - cell n in the Cartesian grid (see k_sys_plan)
        const int R = m.nimg;
        for (int nx = -R; nx <= R; ++nx)
            for (int ny = -R; ny <= R; ++ny)
                for (int nz = -R; nz <= R; ++nz) {
                    const float fx = (float)nx, fy = (float)ny, fz = (float)nz;
                    const double q[3] = {(double)pi.x - ((double)m.H[0] * fx + (double)m.H[1] * fy + (double)m.H[2] * fz),
                                         (double)pi.y - ((double)m.H[3] * fx + (double)m.H[4] * fy + (double)m.H[5] * fz),
                                         (double)pi.z - ((double)m.H[6] * fx + (double)m.H[7] * fy + (double)m.H[8] * fz)};
                    int c[3];
                    bool out = false;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double v = floor((q[d] - (double)m.lo[d]) * (double)m.wsc[d]);
                        if (v < -1.0 || v > (double)m.nc[d]) out = true;          // no cell within one cell of q
                        c[d] = out ? 0 : (int)v;
                        // atoms on the upper face of the bounding box are binned into the last cell (clamped)
                    }
                    if (out) continue;
                    for (int ox = -1; ox <= 1; ++ox) {
                        const int cx = c[0] + ox;
                        if (cx < 0 || cx >= m.nc[0]) continue;
                        for (int oy = -1; oy <= 1; ++oy) {
                            const int cy = c[1] + oy;
                            if (cy < 0 || cy >= m.nc[1]) continue;
                            for (int oz = -1; oz <= 1; ++oz) {
                                const int cz = c[2] + oz;
                                if (cz < 0 || cz >= m.nc[2]) continue;
                                const int cid = m.cell_off + (cx * m.nc[1] + cy) * m.nc[2] + cz;
                                scan_cell<FILL, true>(i, lane, pi, pos, m, cell_start[cid], cell_start[cid + 1], sorted_atoms, cutoff,
                                                      n_owned, fx, fy, fz, row_buf, found);
                            }
                        }
                    }
                }
        return found;
    }
    // Modes 0 / 1: lane k < 27 owns neighbour cell k and reads its atom range; the 27 ranges are then walked as ONE
    // candidate stream with all 32 lanes busy (a cell of the water box holds ~12 atoms: a loop over cells left 20 lanes
    // idle and chained 27 dependent loads per atom - 0.6 ms per pass on the 98,304-atom box).
    int c[3];
    cell_coords(m, pi.x, pi.y, pi.z, c);
    const bool per = m.mode != 0;
    int s0 = 0, cnt = 0;
    if (lane < 27) {
        const int o[3] = {lane / 9 - 1, (lane / 3) % 3 - 1, lane % 3 - 1};
        int cc[3];
        bool ok = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            cc[d] = c[d] + o[d];
            if (per) {      // grids of one or two cells: every cell is visited once (offsets are de-duplicated)
                if ((m.nc[d] == 1 && o[d] != 0) || (m.nc[d] == 2 && o[d] < 0)) ok = false;
                cc[d] = (cc[d] + m.nc[d]) % m.nc[d];
            } else if (cc[d] < 0 || cc[d] >= m.nc[d]) ok = false;
        }
        if (ok) {
            const int cid = m.cell_off + (cc[0] * m.nc[1] + cc[1]) * m.nc[2] + cc[2];
            s0 = cell_start[cid];
            cnt = cell_start[cid + 1] - s0;
        }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        // owner cell of candidate t: the last lane whose exclusive prefix is <= t (prefixes are non-decreasing)
        int k = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int probe = k + step;
            const int e = __shfl_sync(0xffffffffu, excl, probe & 31);
            if (probe < 32 && e <= t) k = probe;
        }
        const int ks0 = __shfl_sync(0xffffffffu, s0, k), kex = __shfl_sync(0xffffffffu, excl, k);
        bool pass = false;
        int j = -1;
        if (t < total) {
            j = sorted_atoms[ks0 + (t - kex)];
            if (j != i && (i < n_owned || j < n_owned)) {   // ghost-ghost pairs belong to other ranks
                float3 d = make_float3(__fsub_rn(pi.x, pos[3 * j]), __fsub_rn(pi.y, pos[3 * j + 1]),
                                       __fsub_rn(pi.z, pos[3 * j + 2]));
                d = nn_min_image(d, m, nullptr);
                pass = nn_norm3(d) < cutoff;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (FILL && pass) {
            const int slot = found + __popc(mask & ((1u << lane) - 1u));
            if (slot < NN_MAX_DEGREE) row_buf[slot] = j;
        }
        found += __popc(mask);
    }
    return found;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_nbr_count(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N,
            const SysMeta* __restrict__ meta, const int* __restrict__ cell_start,
            const int* __restrict__ sorted_atoms, float cutoff, int n_owned, int* __restrict__ deg,
            int* __restrict__ status) {
    int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= N) return;
    if (i >= n_owned) return;      // ghost rows stay empty: every consumer walks the rows of OWNED atoms only (deg is zero-filled)
    SysMeta m = meta[batch[i]];
    int found = visit_neighbours<false>(i, lane, pos, m, cell_start, sorted_atoms, cutoff, n_owned, nullptr);
    if (lane == 0) {
        if (found > NN_MAX_DEGREE) { atomicMax(&status[NN_ST_ROW_OVERFLOW], found); found = NN_MAX_DEGREE; }
        deg[i] = found;
    }
}

__global__ void k_finish_count(const int* __restrict__ row_ptr, int N, int cap_edges, int* __restrict__ status) {
    int E = row_ptr[N];
    status[NN_ST_N_EDGES] = E;
    // on overflow the pair arrays are not (re)written: expose zero pairs so no kernel reads stale entries
    status[NN_ST_N_PAIRS] = 0;               // set by k_finish_pairs once the pair table exists
    if (E > cap_edges) status[NN_ST_EDGE_OVERFLOW] = E;
}
// P = number of forward pairs = E / 2 for a complete (symmetric) list; with empty ghost rows (domain decomposition) an
// owned-ghost pair has ONE directed edge, so P = E_owned-owned / 2 + E_owned-ghost.
__global__ void k_finish_pairs(const int* __restrict__ pair_ptr, int N, int cap_pairs, int* __restrict__ status) {
    const int P = pair_ptr[N];
    if (status[NN_ST_EDGE_OVERFLOW] != 0) return;
    if (P > cap_pairs) { status[NN_ST_EDGE_OVERFLOW] = max(2 * P, status[NN_ST_N_EDGES] + 2); return; }
    status[NN_ST_N_PAIRS] = P;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_nbr_fill(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N,
           const SysMeta* __restrict__ meta, const int* __restrict__ cell_start,
           const int* __restrict__ sorted_atoms, float cutoff, int n_owned, const int* __restrict__ row_ptr,
           int cap_edges, int* __restrict__ col, int* __restrict__ fwd_cnt, const int* __restrict__ status) {
    __shared__ int s_rows[kWarpsPerBlock][NN_MAX_DEGREE];
    if (status[NN_ST_EDGE_OVERFLOW] != 0) return;
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = blockIdx.x * kWarpsPerBlock + w;
    if (i >= N || i >= n_owned) return;      // ghost rows: no edges, no forward pairs (fwd_cnt is zero-filled)
    SysMeta m = meta[batch[i]];
    int* buf = s_rows[w];
    int found = visit_neighbours<true>(i, lane, pos, m, cell_start, sorted_atoms, cutoff, n_owned, buf);
    if (found > NN_MAX_DEGREE) found = NN_MAX_DEGREE;
    __syncwarp();
    const int base = row_ptr[i];
    int fwd = 0;
    // rank sort (rows are short): position of j = number of smaller entries
    for (int k = lane; k < found; k += 32) {
        int j = buf[k];
        int rank = 0;
        for (int t = 0; t < found; ++t) rank += (buf[t] < j);
        if (base + rank < cap_edges) col[base + rank] = j;
        fwd += (j > i);
    }
    for (int o = 16; o > 0; o >>= 1) fwd += __shfl_xor_sync(0xffffffffu, fwd, o);
    if (lane == 0) fwd_cnt[i] = fwd;
}

// Pair table: forward edges (j > i) are the tail of each sorted row and define the pairs; a reversed
// edge (j < i) finds its pair by binary search of i in the tail of row j (the edge set is symmetric).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_pair_build(const float* __restrict__ pos, const int64_t* __restrict__ batch, int N,
             const SysMeta* __restrict__ meta, const int* __restrict__ row_ptr, const int* __restrict__ col,
             const int* __restrict__ pair_ptr, int cap_pairs, int* __restrict__ edge_pair,
             int* __restrict__ pair_i, int* __restrict__ pair_j, float* __restrict__ pair_disp,
             int* __restrict__ status) {
    if (status[NN_ST_EDGE_OVERFLOW] != 0) return;
    int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= N) return;
    SysMeta m = meta[batch[i]];
    const int r0 = row_ptr[i], r1 = row_ptr[i + 1];
    const int nfwd = pair_ptr[i + 1] - pair_ptr[i];
    const int f0 = r1 - nfwd;
    const float3 pi = make_float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    for (int e = r0 + lane; e < r1; e += 32) {
        int j = col[e];
        if (e >= f0) {
            int p = pair_ptr[i] + (e - f0);
            edge_pair[e] = p;
            if (p < cap_pairs) {
                pair_i[p] = i; pair_j[p] = j;
                float3 d = make_float3(__fsub_rn(pi.x, pos[3 * j]), __fsub_rn(pi.y, pos[3 * j + 1]),
                                       __fsub_rn(pi.z, pos[3 * j + 2]));
                d = nn_min_image(d, m, nullptr);
                pair_disp[3 * p] = d.x; pair_disp[3 * p + 1] = d.y; pair_disp[3 * p + 2] = d.z;
            } else {
                atomicMax(&status[NN_ST_EDGE_OVERFLOW], 2 * (p + 1));
            }
        } else {
            int lo = row_ptr[j + 1] - (pair_ptr[j + 1] - pair_ptr[j]), hi = row_ptr[j + 1];
            const int fj = lo;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (col[mid] < i) lo = mid + 1; else hi = mid; }
            int p = -1;
            if (lo < row_ptr[j + 1] && col[lo] == i) p = pair_ptr[j] + (lo - fj);
            else atomicExch(&status[NN_ST_BATCH_UNSORTED], 2);   // asymmetric edge set: cannot happen
            edge_pair[e] = p | (int)0x80000000u;
        }
    }
}

__global__ void k_edge_index(const int* __restrict__ row_ptr, const int* __restrict__ col, int N,
                             long long E, int64_t* __restrict__ edge_index) {
    int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= N) return;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) {
        if (e < E) { edge_index[e] = i; edge_index[E + e] = col[e]; }
    }
}

// rev[e] = index of the reversed edge (j, i) of edge e = (i, j): binary search of i in the (sorted) row of j.  Grouping the
// edges by SOURCE atom is then rev read in row order - the transposed adjacency without a sort (the edge set is symmetric).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_edge_reverse(const int* __restrict__ row_ptr, const int* __restrict__ col, int N, int cap_edges, const int* __restrict__ status,
               int* __restrict__ rev) {
    if (status[NN_ST_EDGE_OVERFLOW] != 0) return;
    int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= N) return;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1] && e < cap_edges; e += 32) {
        const int j = col[e];
        int lo = row_ptr[j], hi = row_ptr[j + 1];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < i) lo = mid + 1; else hi = mid; }
        rev[e] = lo;
    }
}

}  // namespace

const SysMeta* nn_nbr_sysmeta(const nn_nbr* nl) { return (const SysMeta*)nl->workspace; }

extern "C" size_t nn_nbr_workspace_bytes(int32_t n_atoms, int32_t n_systems) {
    return carve(nullptr, 0, n_atoms, n_systems).total;
}

static int check_nbr(const nn_nbr* nl) {
    NN_REQUIRE(nl != nullptr, "null nn_nbr");
    NN_REQUIRE(nl->n_atoms >= 0 && nl->n_systems >= 1, "bad sizes");
    NN_REQUIRE(nl->cap_cells >= 2 * nl->n_atoms + nl->n_systems, "cap_cells < 2*n_atoms + n_systems");
    NN_REQUIRE(nl->workspace_bytes >= nn_nbr_workspace_bytes(nl->n_atoms, nl->n_systems), "workspace too small");
    NN_REQUIRE(nl->pos && nl->cell && nl->batch && nl->sys_ptr && nl->row_ptr && nl->status, "null pointer");
    return 0;
}

extern "C" int nn_nbr_count(const nn_nbr* nl, float cutoff, void* stream) {
    if (int rc = check_nbr(nl)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(NN_STAGE_NBR, s);
    const int N = nl->n_atoms, B = nl->n_systems;
    NbrWs w = carve(nl->workspace, nl->workspace_bytes, N, B);
    const int cap_cells = 2 * N + B + 1;
    cudaMemsetAsync(nl->status, 0, NN_STATUS_WORDS * sizeof(int), s);
    cudaMemsetAsync(w.cell_count, 0, (size_t)(cap_cells + 1) * sizeof(int), s);
    cudaMemsetAsync(w.cell_fill, 0, (size_t)cap_cells * sizeof(int), s);
    cudaMemsetAsync(w.deg, 0, (size_t)(N + 1) * sizeof(int), s);
    cudaMemsetAsync(w.fwd_cnt, 0, (size_t)(N + 1) * sizeof(int), s);
    k_sys_ptr<<<nn_ceil_div(N + 1, 256), 256, 0, s>>>(nl->batch, N, B, nl->sys_ptr, nl->status); NN_LAUNCHED(1);
    k_bounds_init<<<nn_ceil_div(B * 6, 256), 256, 0, s>>>(w.bounds, B); NN_LAUNCHED(1);
    if (N > 0) { k_sys_bounds<<<nn_ceil_div(N, 256), 256, 0, s>>>(nl->pos, nl->batch, N, B, w.bounds); NN_LAUNCHED(1); }
    k_sys_plan<<<1, 1024, 0, s>>>(nl->cell, nl->sys_ptr, w.bounds, B, cutoff, w.meta, nl->status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_nbr_count(plan)");      // checked per launch group so that a failed launch is reported where it happened
    if (N > 0) {
        k_bin_count<<<nn_ceil_div(N, 256), 256, 0, s>>>(nl->pos, nl->batch, N, w.meta, w.atom_cell, w.cell_count); NN_LAUNCHED(1);
        NN_CHECK_LAUNCH("nn_nbr_count(bin)");
        NN_LAUNCHED(exclusive_scan(w.cell_count, w.cell_start, cap_cells + 1, (int*)w.scan_tmp, s));
        k_bin_fill<<<nn_ceil_div(N, 256), 256, 0, s>>>(N, w.atom_cell, w.cell_start, w.cell_fill, w.sorted_atoms); NN_LAUNCHED(1);
        k_nbr_count<<<nn_ceil_div(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
            nl->pos, nl->batch, N, w.meta, w.cell_start, w.sorted_atoms, cutoff, nl->n_owned > 0 ? nl->n_owned : N, w.deg,
            nl->status); NN_LAUNCHED(1);
        NN_CHECK_LAUNCH("nn_nbr_count(count)");
    }
    NN_LAUNCHED(exclusive_scan(w.deg, nl->row_ptr, N + 1, (int*)w.scan_tmp, s));
    k_finish_count<<<1, 1, 0, s>>>(nl->row_ptr, N, nl->cap_edges, nl->status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_nbr_count");
    return 0;
}

extern "C" int nn_nbr_fill(const nn_nbr* nl, float cutoff, void* stream) {
    if (int rc = check_nbr(nl)) return rc;
    NN_REQUIRE(nl->col && nl->edge_pair && nl->pair_ptr && nl->pair_i && nl->pair_j && nl->pair_disp, "null output");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(NN_STAGE_NBR, s);
    const int N = nl->n_atoms, B = nl->n_systems;
    NbrWs w = carve(nl->workspace, nl->workspace_bytes, N, B);
    if (N > 0) {
        k_nbr_fill<<<nn_ceil_div(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
            nl->pos, nl->batch, N, w.meta, w.cell_start, w.sorted_atoms, cutoff, nl->n_owned > 0 ? nl->n_owned : N,
            nl->row_ptr, nl->cap_edges,
            nl->col, w.fwd_cnt, nl->status); NN_LAUNCHED(1);
        NN_CHECK_LAUNCH("nn_nbr_fill(fill)");
    }
    NN_LAUNCHED(exclusive_scan(w.fwd_cnt, nl->pair_ptr, N + 1, (int*)w.scan_tmp, s));
    k_finish_pairs<<<1, 1, 0, s>>>(nl->pair_ptr, N, nl->cap_pairs, nl->status); NN_LAUNCHED(1);
    if (N > 0) {
        k_pair_build<<<nn_ceil_div(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
            nl->pos, nl->batch, N, w.meta, nl->row_ptr, nl->col, nl->pair_ptr, nl->cap_pairs, nl->edge_pair,
            nl->pair_i, nl->pair_j, nl->pair_disp, nl->status); NN_LAUNCHED(1);
    }
    NN_CHECK_LAUNCH("nn_nbr_fill");
    return 0;
}

extern "C" int nn_nbr_edge_reverse(const nn_nbr* nl, int32_t* rev, void* stream) {
    NN_REQUIRE(nl && rev && nl->row_ptr && nl->col, "null pointer");
    if (nl->n_atoms == 0) return 0;
    k_edge_reverse<<<nn_ceil_div(nl->n_atoms, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        nl->row_ptr, nl->col, nl->n_atoms, nl->cap_edges, nl->status, rev); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_nbr_edge_reverse");
    return 0;
}

extern "C" int nn_nbr_edge_index(const nn_nbr* nl, int64_t* edge_index, int64_t n_edges, void* stream) {
    NN_REQUIRE(nl && edge_index, "null pointer");
    if (nl->n_atoms == 0 || n_edges == 0) return 0;
    k_edge_index<<<nn_ceil_div(nl->n_atoms, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        nl->row_ptr, nl->col, nl->n_atoms, n_edges, edge_index); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_nbr_edge_index");
    return 0;
}
