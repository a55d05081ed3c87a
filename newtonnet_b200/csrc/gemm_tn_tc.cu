// Weight-gradient contraction on the tensor cores:  G[128,128] = X[M,128]^T . Y[M,128]   (the reduction runs over ROWS).
//
// Autograd transposes of every nn.Linear of the training step (reference train/trainer.py:303-313 calls
// loss.backward(); the 128 -> 128 linears are models/newtonnet.py:181-199, models/output.py:90-96): dL/dW = X^T dY, and
// again inside the double backward.  M is the number of atoms or of directed edges, so these products were 48 % of the
// config-5 step as an fp32 FFMA kernel (profiles/r1e_c5_launches_summary.txt).
//
// Both UMMA operands are TRANSPOSES of row blocks: for a block of 32 rows, A[i][k] = X[r0 + k][i] and B[n][k] =
// Y[r0 + k][n] are [128 x 32] K-major tiles.  Producer threads own one column and four consecutive rows, so that one
// 16-byte chunk of the 128B-swizzled tile is written per store (coalesced 128-byte row segments on the load side),
// split into tf32 hi / lo parts (3xTF32: lo*hi + hi*lo + hi*hi, as gemm_ts.cu).  Each CTA accumulates its row range in
// one 128-column TMEM accumulator and writes a [128,128] partial; k_tn_reduce sums the partials in fixed order.
#include "tc_common.cuh"

namespace {

using namespace tc;
constexpr int STAGES = 2;
constexpr int PRODUCER_THREADS = 256;
constexpr int MMA_WARP = PRODUCER_THREADS / 32;
constexpr int THREADS = PRODUCER_THREADS + 32;
constexpr uint32_t STAGE_BYTES = 4 * BLK_BYTES;            // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256;
constexpr uint32_t TMEM_COLS = 128;

__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tn_tc(const float* __restrict__ X, const float* __restrict__ Y, int M, int rows_per_cta, float* __restrict__ partial) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sT = base;
    const uint32_t sBar = base + STAGES * STAGE_BYTES;
    const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_done = bar_empty + 8 * STAGES, tmem_slot = bar_done + 8;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, PRODUCER_THREADS); mbar_init(bar_empty + 8 * s, 1); }
            mbar_init(bar_done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    const int r_begin = blockIdx.x * rows_per_cta;
    const int r_end = min(M, r_begin + rows_per_cta);
    const int n_blocks = r_end > r_begin ? (r_end - r_begin + KB - 1) / KB : 0;

    if (warp < MMA_WARP) {
        // ===================== producers: 32-row blocks of X and Y -> transposed hi / lo tiles =====================
        const int col = threadIdx.x & 127, half = threadIdx.x >> 7;          // chunks 4*half .. 4*half+3 (rows 16*half .. +15)
        uint32_t stage = 0, phase = 0;
        float xv[16], yv[16], xn[16], yn[16];
        auto load = [&](int b, float (&xo)[16], float (&yo)[16]) {
            const int r0 = r_begin + b * KB + 16 * half;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int r = r0 + t;
                const bool ok = b < n_blocks && r < r_end;
                xo[t] = ok ? X[(size_t)r * 128 + col] : 0.f;
                yo[t] = ok ? Y[(size_t)r * 128 + col] : 0.f;
            }
        };
        load(0, xv, yv);
        for (int b = 0; b < n_blocks; ++b) {
            load(b + 1, xn, yn);                 // the next block's rows are in flight while this one is split and stored
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            uint8_t* st = smem_gen + (sT - base) + stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t off = swz_offset_bytes(col, 4 * half + j);
                const float4 x = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
                const float4 y = make_float4(yv[4 * j], yv[4 * j + 1], yv[4 * j + 2], yv[4 * j + 3]);
                const float4 xh = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                const float4 yh = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                *reinterpret_cast<float4*>(st + off) = xh;
                *reinterpret_cast<float4*>(st + BLK_BYTES + off) = f4_sub(x, xh);
                *reinterpret_cast<float4*>(st + 2 * BLK_BYTES + off) = yh;
                *reinterpret_cast<float4*>(st + 3 * BLK_BYTES + off) = f4_sub(y, yh);
            }
            fence_proxy_async();
            mbar_arrive(bar_full + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
            for (int t = 0; t < 16; ++t) { xv[t] = xn[t]; yv[t] = yn[t]; }
        }
    } else if (lane == 0) {
        // ===================== MMA issue =====================
        uint32_t stage = 0, phase = 0;
        for (int b = 0; b < n_blocks; ++b) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_hi = sT + stage * STAGE_BYTES, a_lo = a_hi + BLK_BYTES, b_hi = a_hi + 2 * BLK_BYTES, b_lo = a_hi + 3 * BLK_BYTES;
#pragma unroll
            for (int ks = 0; ks < KB / 8; ++ks) {
                const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
                const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
                umma_tf32(tmem_base, dal, dbh, (b | ks) != 0);
                umma_tf32(tmem_base, dah, dbl, 1);
                umma_tf32(tmem_base, dah, dbh, 1);
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_done);
    }
    __syncwarp();
    // ===================== epilogue: accumulator -> this CTA's partial =====================
    if (warp < MMA_WARP) {
        float* P = partial + (size_t)blockIdx.x * 128 * 128;
        const int q = warp & 3, chalf = warp >> 2;
        const int row = q * 32 + lane;
        if (n_blocks > 0) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int c0 = chalf * 64 + c * 32;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st4(P + (size_t)row * 128 + c0 + 4 * j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) st4(P + (size_t)row * 128 + chalf * 64 + 4 * j, f4_zero());
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// out = (accumulate ? out : 0) + sum of the partials, in a fixed order.  The sum is a latency chain of L2 reads, so it is
// spread wide: a block owns 32 consecutive float4 of the 128 x 128 result (one per lane, 512 contiguous bytes per warp
// load) and its eight warps each sum a contiguous eighth of the partials with four independent chains; the eight slice
// sums are then added in slice order through shared memory.  148 partials: 5 dependent load rounds per thread instead
// of 19 (the single-pass version took 16 us per call and was 15 % of the training step, profiles/r2_c5_final_launches*).
constexpr int RED_SLICES = 8;
__global__ void __launch_bounds__(32 * RED_SLICES) k_tn_reduce(const float* __restrict__ partial, int n_partial, float* __restrict__ out,
                                                               int accumulate) {
    __shared__ float4 part[RED_SLICES][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;                       // one float4 of the result
    const int per = (n_partial + RED_SLICES - 1) / RED_SLICES;
    const int b0 = slice * per, b1 = min(n_partial, b0 + per);
    float4 acc[4] = {f4_zero(), f4_zero(), f4_zero(), f4_zero()};
    int b = b0;
    for (; b + 4 <= b1; b += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = f4_add(acc[u], ld4(partial + (size_t)(b + u) * 128 * 128 + 4 * t));
    }
    for (; b < b1; ++b) acc[0] = f4_add(acc[0], ld4(partial + (size_t)b * 128 * 128 + 4 * t));
    part[slice][lane] = f4_add(f4_add(acc[0], acc[1]), f4_add(acc[2], acc[3]));
    __syncthreads();
    if (slice == 0) {
        float4 r = accumulate ? ld4(out + 4 * t) : f4_zero();
#pragma unroll
        for (int s = 0; s < RED_SLICES; ++s) r = f4_add(r, part[s][lane]);
        st4(out + 4 * t, r);
    }
}

bool g_attr = false;

}  // namespace

// number of CTAs (= partials) for m rows: at least 6 row blocks of 32 per CTA, at most one CTA per SM (measured at 30k
// rows: 4 blocks per CTA 16 + 16 us (kernel + reduction), 16 blocks per CTA 29 + 9 us without the register prefetch)
int nn_gemm_tn_tc_ctas(int m) {
    int b = nn_ceil_div(m, 6 * tc::KB);
    const int sms = nn_num_sms();
    return b < 1 ? 1 : (b > sms ? sms : b);
}

int nn_gemm_tn_tc_launch(const float* X, const float* Y, int m, float* out, void* workspace, int accumulate, cudaStream_t s) {
    if (!g_attr) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) { nn_set_error("nn_gemm128_tn(tc): cannot set %u B dynamic smem: %s", SMEM_BYTES, cudaGetErrorString(e)); return -2; }
        g_attr = true;
    }
    const int n = nn_gemm_tn_tc_ctas(m);
    int rows = nn_ceil_div(m, n);
    rows = nn_ceil_div(rows, tc::KB) * tc::KB;
    k_gemm_tn_tc<<<n, THREADS, SMEM_BYTES, s>>>(X, Y, m, rows, (float*)workspace); NN_LAUNCHED(1);
    k_tn_reduce<<<128 * 128 / 4 / 32, 32 * RED_SLICES, 0, s>>>((const float*)workspace, n, out, accumulate); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_gemm128_tn(tc)");
    return 0;
}
