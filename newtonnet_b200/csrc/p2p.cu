// Domain-decomposition communication over NVLink peer memory (no NCCL call on the data path).
//
// Every rank owns one arena of plain cudaMalloc memory (so that CUDA IPC handles work whatever allocator the
// host framework uses) that its peers map: landing buffers for ghost feature rows, flag words, the complete
// force array and a table of per-rank partial sums.  All step-dependent state (the step counter that the flag
// epochs derive from) lives in DEVICE memory, so a whole decomposed evaluation - neighbour rebuild, phases,
// halo exchanges, final reduction - is captured once as a CUDA graph and replayed with one launch per step.
//
//   k_dd_begin      local positions = pos[l2g]; "has an owned atom moved more than skin/2" flag; ++step
//   k_halo_push     packs the rows this rank owes its peers and stores them straight into the peers' landing
//                   buffers (st.global through NVSwitch), fences, last block raises flag[my_rank] = epoch in
//                   every peer
//   k_halo_wait_copy  spins (bounded) on the local flags, then copies the landing buffer into the ghost tail
//   k_dd_push_results owned forces -> every rank's complete force array (owner-only writes, no all-reduce),
//                   partial energy / virial / stress / status -> every rank's partial table, flags
//   k_dd_finish     waits, sums the partials in rank order (fixed order: identical bits on every rank)
//
// Two landing buffers per channel alternate by epoch parity: a peer can run at most one exchange ahead, because its
// exchange e+2 needs data this rank only sends after it has consumed exchange e.  Two independent channels
// (own landing buffers, flags, epoch sequence) let a second stream exchange rows that are not on the critical
// path (f_out of the previous layer, abar) while the main stream computes.
#include <string.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4_cg(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

struct PushPeer {
    float* landing[2];       // peer's landing buffers of this channel (mapped into this process)
    int* flags;              // peer's flag array of this channel, one int per source rank
    int row_offset;          // first row of this rank's block inside the peer's ghost order
    int send_begin, send_end;   // range of this peer inside send_idx
};
struct PushArgs {
    PushPeer peer[NN_DD_MAX_RANKS];
    int n_peers, my_rank, stride, seq;
};

__device__ __forceinline__ int dd_epoch(const int* step, int stride, int seq) { return step[0] * stride + seq; }

// grid-stride over all (row, float4) items of all peers; the last block to finish publishes the flags
__global__ void k_halo_push(const float* __restrict__ src, const int* __restrict__ send_idx, int width4,
                            PushArgs a, const int* __restrict__ step, unsigned int* __restrict__ done_counter) {
    const int epoch = dd_epoch(step, a.stride, a.seq), par = epoch & 1;
    const int total_rows = a.peer[a.n_peers - 1].send_end;
    const long long total = (long long)total_rows * width4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / width4), c = (int)(t % width4);
        int p = 0;
        while (k >= a.peer[p].send_end) ++p;
        const PushPeer& pp = a.peer[p];
        const float4 v = ld4(src + ((size_t)send_idx[k] * width4 + c) * 4);
        st4(pp.landing[par] + ((size_t)(pp.row_offset + k - pp.send_begin) * width4 + c) * 4, v);
    }
    __threadfence_system();                      // this thread's peer stores are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {             // every block has fenced its stores
            *done_counter = 0;
            __threadfence_system();
            for (int p = 0; p < a.n_peers; ++p) {        // every peer, also those that get no rows: keeps ranks in lockstep
                volatile int* f = a.peer[p].flags + a.my_rank;
                *f = epoch;
            }
            __threadfence_system();
        }
    }
}

// every block waits until all peers have published `epoch` (gives up after ~2 s and records the failure), then the
// grid copies the landing buffer of this epoch's parity into the ghost tail
__global__ void k_halo_wait_copy(volatile int* flags, int world, int my_rank, const int* __restrict__ step, int stride, int seq,
                                 const float* __restrict__ landing0, const float* __restrict__ landing1,
                                 float* __restrict__ dst, long long n4, int* __restrict__ status) {
    const int epoch = dd_epoch(step, stride, seq);
    if ((int)threadIdx.x < world && (int)threadIdx.x != my_rank) {
        long long spins = 0;
        while (flags[threadIdx.x] < epoch) {
            __nanosleep(100);
            if (++spins > 20000000LL) { atomicExch(&status[NN_DD_ST_TIMEOUT], 1 + (int)threadIdx.x); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    const float* src = (epoch & 1) ? landing1 : landing0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x)
        st4(dst + 4 * t, ld4_cg(src + 4 * t));
}

// local positions (owned, then ghosts) gathered from the replicated position array; sticky "stale plan" flag when an
// OWNED atom is further than skin/2 from where it was when the plan was made (every atom is owned by exactly one rank
// and the flags are OR-ed across ranks by k_dd_finish); block 0 advances the step counter and clears the step status.
__global__ void k_dd_begin(const float* __restrict__ pos, const float* __restrict__ pos_ref, const float* __restrict__ cell,
                           const float* __restrict__ cell_ref, const int64_t* __restrict__ z, const int* __restrict__ l2g,
                           int n_local, int n_owned, float half_skin2, float* __restrict__ pos_local,
                           int64_t* __restrict__ z_local, int* __restrict__ step, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        step[0] += 1; status[NN_DD_ST_TIMEOUT] = 0;
        bool same = true;
        for (int k = 0; k < 9; ++k) same = same && cell[k] == cell_ref[k];      // the bricks were cut for cell_ref
        if (!same) status[NN_DD_ST_STALE] = 1;
    }
    if (i >= n_local) return;
    const int g = l2g[i];
    const float x = pos[3 * g], y = pos[3 * g + 1], zz = pos[3 * g + 2];
    pos_local[3 * i] = x; pos_local[3 * i + 1] = y; pos_local[3 * i + 2] = zz;
    z_local[i] = z[g];
    if (i < n_owned) {
        const float dx = x - pos_ref[3 * g], dy = y - pos_ref[3 * g + 1], dz = zz - pos_ref[3 * g + 2];
        if (dx * dx + dy * dy + dz * dz > half_skin2 || !(x == x)) status[NN_DD_ST_STALE] = 1;
    }
}

struct ResArgs {
    float* forces_full[NN_DD_MAX_RANKS];     // every rank's complete force array (index = rank, own included)
    float* partials[NN_DD_MAX_RANKS];        // every rank's partial table [world][NN_DD_PARTIAL]
    int* flags[NN_DD_MAX_RANKS];             // channel-0 flags of every rank
    int world, my_rank, stride, seq;
};

__global__ void k_dd_push_results(const float* __restrict__ forces_owned, const int* __restrict__ l2g, int n_owned,
                                  const float* __restrict__ energy, const float* __restrict__ virial,
                                  const float* __restrict__ stress, const int* __restrict__ nbr_status,
                                  const int* __restrict__ dd_status, ResArgs a, const int* __restrict__ step,
                                  unsigned int* __restrict__ done_counter) {
    const int epoch = dd_epoch(step, a.stride, a.seq);
    const long long total = (long long)n_owned * a.world;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / n_owned), i = (int)(t % n_owned);
        float* dst = a.forces_full[r] + 3 * (size_t)l2g[i];
        dst[0] = forces_owned[3 * i]; dst[1] = forces_owned[3 * i + 1]; dst[2] = forces_owned[3 * i + 2];
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
        float* dst = a.partials[threadIdx.x] + (size_t)a.my_rank * NN_DD_PARTIAL;
        dst[0] = energy[0];
        for (int k = 0; k < 9; ++k) { dst[1 + k] = virial ? virial[k] : 0.f; dst[10 + k] = stress ? stress[k] : 0.f; }
        dst[19] = nbr_status[NN_ST_EDGE_OVERFLOW] != 0 ? (float)nbr_status[NN_ST_EDGE_OVERFLOW] : 0.f;
        dst[20] = (nbr_status[NN_ST_ROW_OVERFLOW] != 0 || nbr_status[NN_ST_BATCH_UNSORTED] != 0 || nbr_status[NN_ST_SINGULAR_CELL] != 0) ? 1.f : 0.f;
        dst[21] = dd_status[NN_DD_ST_STALE] != 0 ? 1.f : 0.f;
        dst[22] = dd_status[NN_DD_ST_TIMEOUT] != 0 ? 1.f : 0.f;
        dst[23] = (float)nbr_status[NN_ST_N_EDGES];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *done_counter = 0;
            __threadfence_system();
            for (int r = 0; r < a.world; ++r)
                if (r != a.my_rank) { volatile int* f = a.flags[r] + a.my_rank; *f = epoch; }
            __threadfence_system();
        }
    }
}

// out_small: energy, virial[9], stress[9] as fp32 (fp64 sums over ranks in rank order); out_status[NN_DD_STATUS_WORDS]
__global__ void k_dd_finish(volatile int* flags, int world, int my_rank, const int* __restrict__ step, int stride, int seq,
                            const float* __restrict__ partials, const float* __restrict__ forces_full, long long n3,
                            float* __restrict__ forces_out, float* __restrict__ out_small, int* __restrict__ dd_status,
                            int* __restrict__ out_status) {
    const int epoch = dd_epoch(step, stride, seq);
    if ((int)threadIdx.x < world && (int)threadIdx.x != my_rank) {
        long long spins = 0;
        while (flags[threadIdx.x] < epoch) {
            __nanosleep(100);
            if (++spins > 20000000LL) { atomicExch(&dd_status[NN_DD_ST_TIMEOUT], 1 + (int)threadIdx.x); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (threadIdx.x < 19) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) s += (double)__ldcg(partials + (size_t)r * NN_DD_PARTIAL + threadIdx.x);
            out_small[threadIdx.x] = (float)s;
        }
        if (threadIdx.x == 32) {
            float over = 0.f, bad = 0.f, stale = 0.f, tmo = dd_status[NN_DD_ST_TIMEOUT] != 0 ? 1.f : 0.f, edges = 0.f;
            for (int r = 0; r < world; ++r) {
                const float* p = partials + (size_t)r * NN_DD_PARTIAL;
                over = fmaxf(over, __ldcg(p + 19)); bad = fmaxf(bad, __ldcg(p + 20)); stale = fmaxf(stale, __ldcg(p + 21));
                tmo = fmaxf(tmo, __ldcg(p + 22)); edges += __ldcg(p + 23);
            }
            out_status[NN_DD_ST_STALE] = stale != 0.f;
            out_status[NN_DD_ST_TIMEOUT] = tmo != 0.f;
            out_status[NN_DD_ST_OVERFLOW] = over != 0.f;
            out_status[NN_DD_ST_BAD_INPUT] = bad != 0.f;
            out_status[NN_DD_ST_STEP] = step[0];
            out_status[NN_DD_ST_EDGES] = (int)(edges * (1.0f / 1024.f));     // total directed edges / 1024 (diagnostic)
        }
    }
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n3; t += (long long)gridDim.x * blockDim.x)
        forces_out[t] = __ldcg(forces_full + t);
}

int grid_for(long long items, int threads, int cap) {
    long long g = (items + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int check_comm(const nn_dd_comm* c, int channel) {
    NN_REQUIRE(c != nullptr, "null nn_dd_comm");
    NN_REQUIRE(c->world >= 2 && c->world <= NN_DD_MAX_RANKS, "2..16 ranks");
    NN_REQUIRE(c->rank >= 0 && c->rank < c->world, "bad rank");
    NN_REQUIRE(channel >= 0 && channel < NN_DD_CHANNELS, "bad channel");
    NN_REQUIRE(c->step && c->done && c->status && c->flags[channel], "null device pointer");
    return 0;
}

}  // namespace

extern "C" int nn_p2p_alloc(size_t bytes, void** ptr) {
    NN_REQUIRE(ptr, "null pointer");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) { nn_set_error("nn_p2p_alloc: %s", cudaGetErrorString(e)); return -2; }
    cudaMemset(*ptr, 0, bytes);
    return 0;
}
extern "C" int nn_p2p_free(void* ptr) { cudaFree(ptr); return 0; }
extern "C" int nn_p2p_get_handle(void* ptr, void* handle64) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { nn_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return -2; }
    memcpy(handle64, &h, sizeof(h));
    return 0;
}
extern "C" int nn_p2p_open_handle(const void* handle64, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { nn_set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}
extern "C" int nn_p2p_close_handle(void* ptr) { cudaIpcCloseMemHandle(ptr); return 0; }

extern "C" int nn_dd_begin(const nn_dd_comm* c, const float* pos, const float* pos_ref, const float* cell, const float* cell_ref,
                           const int64_t* z, const int32_t* l2g, int32_t n_local, float skin, float* pos_local,
                           int64_t* z_local, void* stream) {
    NN_REQUIRE(c && pos && pos_ref && cell && cell_ref && z && l2g && pos_local && z_local && c->step && c->status, "null pointer");
    const float h = 0.5f * skin;
    k_dd_begin<<<grid_for(n_local > 0 ? n_local : 1, 256, 1 << 20), 256, 0, (cudaStream_t)stream>>>(
        pos, pos_ref, cell, cell_ref, z, l2g, n_local, c->n_owned, h * h, pos_local, z_local, c->step, c->status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_dd_begin");
    return 0;
}

extern "C" int nn_dd_halo_push(const nn_dd_comm* c, int32_t channel, int32_t seq, const float* rows, int32_t width, void* stream) {
    if (int rc = check_comm(c, channel)) return rc;
    NN_REQUIRE(width > 0 && width % 4 == 0 && width <= NN_DD_MAX_WIDTH, "width must be a multiple of 4, at most 384");
    NN_REQUIRE(seq >= 0 && seq < c->stride[channel], "seq out of range");
    PushArgs a;
    a.n_peers = 0; a.my_rank = c->rank; a.stride = c->stride[channel]; a.seq = seq;
    int last_end = 0;
    for (int s = 0; s < c->world; ++s) {
        if (s == c->rank) continue;
        NN_REQUIRE(c->send_begin[s] == last_end && c->send_end[s] >= c->send_begin[s], "send ranges must be contiguous in rank order");
        last_end = c->send_end[s];
        PushPeer& p = a.peer[a.n_peers++];
        p.landing[0] = c->peer_landing[channel][0][s]; p.landing[1] = c->peer_landing[channel][1][s];
        p.flags = c->peer_flags[channel][s];
        p.row_offset = c->row_offset[s]; p.send_begin = c->send_begin[s]; p.send_end = c->send_end[s];
        NN_REQUIRE(p.flags && (p.send_end == p.send_begin || (p.landing[0] && p.landing[1])), "peer memory not mapped");
    }
    const long long total = (long long)last_end * (width / 4);
    k_halo_push<<<grid_for(total, 256, nn_num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(rows, c->send_idx, width / 4, a, c->step,
                                                                                 c->done + channel); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_dd_halo_push");
    return 0;
}

extern "C" int nn_dd_halo_wait(const nn_dd_comm* c, int32_t channel, int32_t seq, float* ghost_rows, int32_t width, void* stream) {
    if (int rc = check_comm(c, channel)) return rc;
    NN_REQUIRE(width > 0 && width % 4 == 0 && width <= NN_DD_MAX_WIDTH, "width must be a multiple of 4, at most 384");
    const long long n4 = (long long)c->n_ghost * (width / 4);
    k_halo_wait_copy<<<grid_for(n4, 256, nn_num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(
        c->flags[channel], c->world, c->rank, c->step, c->stride[channel], seq, c->landing[channel][0], c->landing[channel][1],
        ghost_rows, n4, c->status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_dd_halo_wait");
    return 0;
}

extern "C" int nn_dd_finish(const nn_dd_comm* c, int32_t seq, const float* forces_owned, const int32_t* l2g, const float* energy,
                            const float* virial, const float* stress, const int32_t* nbr_status, float* forces_out,
                            float* out_small, int32_t* out_status, void* stream) {
    if (int rc = check_comm(c, 0)) return rc;
    NN_REQUIRE(forces_owned && l2g && energy && nbr_status && forces_out && out_small && out_status, "null pointer");
    NN_REQUIRE(c->forces_full && c->partials, "null arena pointer");
    ResArgs a;
    a.world = c->world; a.my_rank = c->rank; a.stride = c->stride[0]; a.seq = seq;
    for (int r = 0; r < c->world; ++r) {
        a.forces_full[r] = r == c->rank ? c->forces_full : c->peer_forces_full[r];
        a.partials[r] = r == c->rank ? c->partials : c->peer_partials[r];
        a.flags[r] = r == c->rank ? c->flags[0] : c->peer_flags[0][r];
        NN_REQUIRE(a.forces_full[r] && a.partials[r] && a.flags[r], "peer memory not mapped");
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)c->n_owned * c->world;
    k_dd_push_results<<<grid_for(total, 256, nn_num_sms() * 2), 256, 0, s>>>(forces_owned, l2g, c->n_owned, energy, virial, stress,
                                                                    nbr_status, c->status, a, c->step, c->done); NN_LAUNCHED(1);
    const long long n3 = (long long)c->n_atoms_total * 3;
    k_dd_finish<<<grid_for(n3, 256, nn_num_sms() * 2), 256, 0, s>>>(c->flags[0], c->world, c->rank, c->step, c->stride[0], seq, c->partials,
                                                           c->forces_full, n3, forces_out, out_small, c->status, out_status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_dd_finish");
    return 0;
}
