// Primitives of the differentiable (training) path - SURVEY.md section 8a row T.
//
// The training step needs d(loss)/d(parameters) through the forces (double backward, reference
// train/trainer.py:303-313 + models/output.py:66-73 with create_graph=True).  Instead of hand-deriving a
// second-order sweep, the training forward is composed from a closed set of primitives whose
// derivatives are again these primitives, so autograd can differentiate twice:
//   X @ B            nn_gemm128                  d/dX = dY @ B^T (same op), d/dB = X^T dY (nn_gemm128_tn)
//   X^T Y            nn_gemm128_tn               d/dX = Y @ G^T,  d/dY = X @ G       (nn_gemm128)
//   gather rows      nn_halo_pack (out = src[idx])   d/dsrc = segment sum
//   segment sum      nn_segment_sum              d/dsrc = gather rows
// nn_segment_sum is the deterministic replacement of torch_geometric.utils.scatter(reduce='sum')
// (reference call sites models/newtonnet.py:214,226, models/output.py:246).
#include "common.cuh"

namespace {

// out[i, :] = sum_{k in [row_ptr[i], row_ptr[i+1])} src[perm ? perm[k] : k, :]   (fixed order)
__global__ void k_segment_sum(const float* __restrict__ src, const int* __restrict__ perm,
                              const int* __restrict__ row_ptr, int n_rows, int width4, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_rows) return;
    const int r0 = row_ptr[i], r1 = row_ptr[i + 1];
    for (int c = lane; c < width4; c += 32) {
        float4 acc = f4_zero();
        for (int k = r0; k < r1; ++k) {
            const int e = perm ? perm[k] : k;
            acc = f4_add(acc, ld4(src + ((size_t)e * width4 + c) * 4));
        }
        st4(out + ((size_t)i * width4 + c) * 4, acc);
    }
}

// partial[b] = X[rows of block b]^T @ Y[rows of block b]   (128 x 128), 256 threads, 8x8 outputs per thread
constexpr int TN_ROWS = 16;
__global__ void __launch_bounds__(256, 2)
k_gemm_tn_partial(const float* __restrict__ X, const float* __restrict__ Y, int M, int rows_per_block,
                  float* __restrict__ partial) {
    __shared__ __align__(16) float sx[TN_ROWS][128];
    __shared__ __align__(16) float sy[TN_ROWS][128];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(M, r_begin + rows_per_block);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int r0 = r_begin; r0 < r_end; r0 += TN_ROWS) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {                      // 16 rows x 32 float4 = 512 float4 per operand
            const int idx = tid + q * 256, row = idx >> 5, c = (idx & 31) * 4;
            const bool ok = r0 + row < r_end;
            st4(&sx[row][c], ok ? ld4(X + (size_t)(r0 + row) * 128 + c) : f4_zero());
            st4(&sy[row][c], ok ? ld4(Y + (size_t)(r0 + row) * 128 + c) : f4_zero());
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TN_ROWS; ++r) {
            const float4 a0 = ld4(&sx[r][ty * 4]), a1 = ld4(&sx[r][64 + ty * 4]);
            const float4 b0 = ld4(&sy[r][tx * 4]), b1 = ld4(&sy[r][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* P = partial + (size_t)blockIdx.x * 128 * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int a = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
        st4(P + (size_t)a * 128 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        st4(P + (size_t)a * 128 + 64 + tx * 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
}

__global__ void k_gemm_tn_reduce(const float* __restrict__ partial, int n_partial, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of the 128 x 128 result
    if (t >= 128 * 128 / 4) return;
    // four independent chains keep four loads in flight; the order of the additions is fixed (deterministic)
    float4 acc[4] = {f4_zero(), f4_zero(), f4_zero(), f4_zero()};
    int b = 0;
    for (; b + 4 <= n_partial; b += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = f4_add(acc[u], ld4(partial + (size_t)(b + u) * 128 * 128 + 4 * t));
    }
    for (; b < n_partial; ++b) acc[0] = f4_add(acc[0], ld4(partial + (size_t)b * 128 * 128 + 4 * t));
    st4(out + 4 * t, f4_add(f4_add(acc[0], acc[1]), f4_add(acc[2], acc[3])));
}

// 256 rows per block until two blocks per SM are reached: a training batch has 3e4 - 5e4 edge rows, and with 1024 rows
// per block only ~40 of the 148 SMs worked (the kernel was 48 % of the c5 training step, ncu launch list)
int tn_blocks(int M) {
    int b = nn_ceil_div(M, 16 * TN_ROWS);
    return b < 1 ? 1 : (b > 296 ? 296 : b);
}

}  // namespace

extern "C" int nn_segment_sum(const float* src, const int32_t* perm, const int32_t* row_ptr, int32_t n_rows,
                              int32_t width, float* out, void* stream) {
    NN_REQUIRE(src && row_ptr && out, "null pointer");
    NN_REQUIRE(width > 0 && width % 4 == 0, "width must be a positive multiple of 4");
    if (n_rows <= 0) return 0;
    k_segment_sum<<<nn_ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(src, perm, row_ptr, n_rows, width / 4, out); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_segment_sum");
    return 0;
}

int nn_gemm_tn_tc_ctas(int m);
int nn_gemm_tn_tc_launch(const float* X, const float* Y, int m, float* out, void* workspace, cudaStream_t s);
extern "C" int nn_get_gemm_backend(void);

// workspace = one [128,128] partial per block / CTA of whichever back-end runs (the larger of the two counts)
extern "C" size_t nn_gemm128_tn_workspace_bytes(int32_t m) {
    const int a = tn_blocks(m), b = nn_gemm_tn_tc_ctas(m);
    return (size_t)(a > b ? a : b) * 128 * 128 * sizeof(float);
}

extern "C" int nn_gemm128_tn(const float* X, const float* Y, int32_t m, float* out, void* workspace, void* stream) {
    NN_REQUIRE(X && Y && out && workspace, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (m <= 0) { cudaMemsetAsync(out, 0, 128 * 128 * sizeof(float), s); return 0; }
    if (nn_get_gemm_backend() >= 1) return nn_gemm_tn_tc_launch(X, Y, m, out, workspace, s);     // tcgen05 3xTF32 (gemm_tn_tc.cu)
    const int nb = tn_blocks(m);
    int rows_per_block = nn_ceil_div(m, nb);
    rows_per_block = nn_ceil_div(rows_per_block, TN_ROWS) * TN_ROWS;
    k_gemm_tn_partial<<<nb, 256, 0, s>>>(X, Y, m, rows_per_block, (float*)workspace); NN_LAUNCHED(1);
    k_gemm_tn_reduce<<<128 * 128 / 4 / 256, 256, 0, s>>>((const float*)workspace, nb, out); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_gemm128_tn");
    return 0;
}
