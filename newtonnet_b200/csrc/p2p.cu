// Halo exchange over NVLink peer memory: ONE kernel packs the rows a rank owes its peers and stores them
// straight into the peers' landing buffers (P2P stores through NVSwitch), then raises a per-source flag
// in the peer's memory; the receiver spins on its own flags.  No NCCL call, no staging copy on the
// sender, transfer overlapped with the pack itself.
//
// Buffers are plain cudaMalloc allocations made here (so that CUDA IPC handles work whatever allocator
// the host framework uses); handles travel between the processes through the caller (torch.distributed
// all_gather of 64-byte blobs).  Two landing buffers alternate by epoch parity: a peer can run at most one
// exchange ahead, because its exchange k+2 needs data this rank only sends after consuming epoch k.
#include <string.h>
#include "common.cuh"

namespace {

struct PushPeer {
    float* landing;          // peer's landing buffer (mapped into this process), [n_ghost_of_peer, width_max]
    int* flags;              // peer's flag array, one int per source rank
    int row_offset;          // first row of this rank's block inside the peer's ghost order
    int send_begin, send_end;   // range of this peer inside send_idx
};

constexpr int kMaxPeers = 16;
struct PushArgs {
    PushPeer peer[kMaxPeers];
    int n_peers;
    int my_rank;
    int epoch;
};

// grid-stride over all (row, float4) items of all peers; the last block to finish publishes the flags
__global__ void k_halo_push(const float* __restrict__ src, const int* __restrict__ send_idx, int width4,
                            PushArgs a, unsigned int* __restrict__ done_counter) {
    const int total_rows = a.peer[a.n_peers - 1].send_end;
    const long long total = (long long)total_rows * width4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / width4), c = (int)(t % width4);
        int p = 0;
        while (k >= a.peer[p].send_end) ++p;
        const PushPeer& pp = a.peer[p];
        const float4 v = ld4(src + ((size_t)send_idx[k] * width4 + c) * 4);
        st4(pp.landing + ((size_t)(pp.row_offset + k - pp.send_begin) * width4 + c) * 4, v);
    }
    __threadfence_system();                      // this thread's peer stores are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {             // every block has fenced its stores
            *done_counter = 0;
            __threadfence_system();
            for (int p = 0; p < a.n_peers; ++p)
                if (a.peer[p].send_end > a.peer[p].send_begin) {
                    volatile int* f = a.peer[p].flags + a.my_rank;
                    *f = a.epoch;
                }
            __threadfence_system();
        }
    }
}

// wait until every expected source has published `epoch`; gives up after ~2 s and records the failure
__global__ void k_halo_wait(volatile int* flags, const int* __restrict__ expect, int world, int epoch,
                            int* __restrict__ status) {
    const int s = threadIdx.x;
    if (s >= world || !expect[s]) return;
    long long spins = 0;
    while (flags[s] < epoch) {
        __nanosleep(200);
        if (++spins > 10000000LL) { atomicExch(status, 1 + s); return; }
    }
    __threadfence_system();
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

// send_idx [n_send] rows of `src` ([*, width]) ordered by destination peer; peer p receives rows
// [send_begin[p], send_end[p]) at row_offset[p] of landing[p]; done_counter is one zeroed uint on this device.
extern "C" int nn_halo_push(const float* src, const int32_t* send_idx, int32_t width, int32_t n_peers,
                            float* const* landing, int32_t* const* flags, const int32_t* row_offset,
                            const int32_t* send_begin, const int32_t* send_end, int32_t my_rank, int32_t epoch,
                            uint32_t* done_counter, void* stream) {
    NN_REQUIRE(width > 0 && width % 4 == 0, "width must be a positive multiple of 4");
    NN_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "1..16 peers");
    PushArgs a;
    a.n_peers = n_peers; a.my_rank = my_rank; a.epoch = epoch;
    for (int p = 0; p < n_peers; ++p)
        a.peer[p] = PushPeer{landing[p], flags[p], row_offset[p], send_begin[p], send_end[p]};
    const long long total = (long long)send_end[n_peers - 1] * (width / 4);
    int grid = (int)((total + 255) / 256);
    grid = grid < 1 ? 1 : (grid > 148 * 4 ? 148 * 4 : grid);
    k_halo_push<<<grid, 256, 0, (cudaStream_t)stream>>>(src, send_idx, width / 4, a, done_counter); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_halo_push");
    return 0;
}

extern "C" int nn_copy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { nn_set_error("nn_copy_d2d: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

extern "C" int nn_halo_wait(int32_t* flags, const int32_t* expect, int32_t world, int32_t epoch, int32_t* status,
                            void* stream) {
    NN_REQUIRE(world >= 1 && world <= 32, "1..32 ranks");
    k_halo_wait<<<1, 32, 0, (cudaStream_t)stream>>>(flags, expect, world, epoch, status); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_halo_wait");
    return 0;
}
