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

__global__ void k_gemm_tn_reduce(const float* __restrict__ partial, int n_partial, float* __restrict__ out, int accumulate) {
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
    const float4 r = f4_add(f4_add(acc[0], acc[1]), f4_add(acc[2], acc[3]));
    st4(out + 4 * t, accumulate ? f4_add(ld4(out + 4 * t), r) : r);
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
int nn_gemm_tn_tc_launch(const float* X, const float* Y, int m, float* out, void* workspace, int accumulate, cudaStream_t s);
extern "C" int nn_get_gemm_backend(void);

// workspace = one [128,128] partial per block / CTA of whichever back-end runs (the larger of the two counts)
extern "C" size_t nn_gemm128_tn_workspace_bytes(int32_t m) {
    const int a = tn_blocks(m), b = nn_gemm_tn_tc_ctas(m);
    return (size_t)(a > b ? a : b) * 128 * 128 * sizeof(float);
}

extern "C" int nn_gemm128_tn_acc(const float* X, const float* Y, int32_t m, float* out, void* workspace, int32_t accumulate,
                                 void* stream) {
    NN_REQUIRE(X && Y && out && workspace, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (m <= 0) { if (!accumulate) cudaMemsetAsync(out, 0, 128 * 128 * sizeof(float), s); return 0; }
    if (nn_get_gemm_backend() >= 1) return nn_gemm_tn_tc_launch(X, Y, m, out, workspace, accumulate, s);     // tcgen05 3xTF32 (gemm_tn_tc.cu)
    const int nb = tn_blocks(m);
    int rows_per_block = nn_ceil_div(m, nb);
    rows_per_block = nn_ceil_div(rows_per_block, TN_ROWS) * TN_ROWS;
    k_gemm_tn_partial<<<nb, 256, 0, s>>>(X, Y, m, rows_per_block, (float*)workspace); NN_LAUNCHED(1);
    k_gemm_tn_reduce<<<128 * 128 / 4 / 256, 256, 0, s>>>((const float*)workspace, nb, out, accumulate); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_gemm128_tn");
    return 0;
}

extern "C" int nn_gemm128_tn(const float* X, const float* Y, int32_t m, float* out, void* workspace, void* stream) {
    return nn_gemm128_tn_acc(X, Y, m, out, workspace, 0, stream);
}

// ---------------------------------------------------------------------------- fused row products of the training path
// The element-wise glue of InteractionNet.forward (reference models/newtonnet.py:211,219-226,231) as six bilinear /
// trilinear row kernels that are CLOSED under differentiation - the gradient of each is a combination of the others - so
// autograd composes forward, backward and double backward from them without falling back to broadcasting ATen kernels
// (which were ~700 of the ~1400 launches of a config-5 training step).  Rows of F = 128 floats, one warp per row,
// float4 per lane; `x3` tensors are [n, 3, F], `u` tensors [n, 3].
//   mul3      out = a * b * c                                   d/da = mul3(g, b, c) ...
//   outer     out[e,c,:] = x[e,:] * u[e,c]                      d/dx = contract_c(g, u),  d/du = row_dot(g, x)
//   contract  out[e,:]   = sum_c x3[e,c,:] * u[e,c]             d/dx3 = outer(g, u),      d/du = row_dot(x3, g)
//   row_dot   out[e,c]   = <x3[e,c,:], x[e,:]>                  d/dx3 = outer(x, g),      d/dx = contract_c(x3, g)
//   mul_b     out[e,c,:] = x[e,:] * y3[e,c,:]                   d/dx = sum_mul_c(g, y3),  d/dy3 = mul_b(x, g)
//   sum_mul_c out[e,:]   = sum_c x3[e,c,:] * y3[e,c,:]          d/dx3 = mul_b(g, y3),     d/dy3 = mul_b(g, x3)
namespace {

__global__ void k_ew_mul3(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                          float* __restrict__ out, long long n4) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x)
        st4(out + 4 * t, f4_mul(ld4(a + 4 * t), f4_mul(ld4(b + 4 * t), ld4(c + 4 * t))));
}

// mode 0 outer, 1 contract_c, 2 row_dot, 3 mul_b, 4 sum_mul_c; p = [n,F] operand, q3 = [n,3,F] operand, u = [n,3] operand
template <int MODE>
__global__ void __launch_bounds__(256)
k_ew_rows(const float* __restrict__ p, const float* __restrict__ q3, const float* __restrict__ u, float* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const size_t r = (size_t)e * kF + 4 * lane, r3 = (size_t)e * 3 * kF + 4 * lane;
    if (MODE == 0) {            // outer(p, u)
        const float4 x = ld4(p + r);
#pragma unroll
        for (int c = 0; c < 3; ++c) st4(out + r3 + c * kF, f4_fma(u[3 * e + c], x, f4_zero()));
    } else if (MODE == 1) {     // contract_c(q3, u)
        float4 acc = f4_zero();
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = f4_fma(u[3 * e + c], ld4(q3 + r3 + c * kF), acc);
        st4(out + r, acc);
    } else if (MODE == 2) {     // row_dot(q3, p)
        const float4 x = ld4(p + r);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = warp_sum(f4_dot(ld4(q3 + r3 + c * kF), x));
            if (lane == 0) out[3 * e + c] = s;
        }
    } else if (MODE == 3) {     // mul_b(p, q3)
        const float4 x = ld4(p + r);
#pragma unroll
        for (int c = 0; c < 3; ++c) st4(out + r3 + c * kF, f4_mul(x, ld4(q3 + r3 + c * kF)));
    } else {                    // sum_mul_c(q3, second [n,3,F] operand passed in p)
        float4 acc = f4_zero();
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = f4_fma(ld4(q3 + r3 + c * kF), ld4(p + r3 + c * kF), acc);
        st4(out + r, acc);
    }
}

// Row product with GATHERED operands: the gathers mn[idx[e]] that feed the message of models/newtonnet.py:211 read straight
// from the (L2-resident, N x 512 B) node table instead of materialising [E,F] copies first.
//   k_ew_gmul:      out[e,:] = a[e,:] * (b ? b[e,:] : 1) * r1[i1[e],:] * (r2 ? r2[i2[e],:] : 1)
__global__ void __launch_bounds__(256)
k_ew_gmul(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ r1, const int* __restrict__ i1,
          const float* __restrict__ r2, const int* __restrict__ i2, float* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const size_t r = (size_t)e * kF + 4 * lane;
    float4 v = f4_mul(ld4(a + r), ld4(r1 + (size_t)i1[e] * kF + 4 * lane));
    if (b) v = f4_mul(v, ld4(b + r));
    if (r2) v = f4_mul(v, ld4(r2 + (size_t)i2[e] * kF + 4 * lane));
    st4(out + r, v);
}

// Equivariant aggregation without [E,3,F] tensors (reference models/newtonnet.py:219-226: delta f_i = sum_{e->i} e1_e u_e +
// e2_e * f_j).  Five kernels that are closed under differentiation; rows3 is an [N,3,F] node table read
// through an edge index, segments are rows of a CSR (perm = edge order within the rows, or null):
//   k_seg_outer:   out[k,c,:] = sum_{e in seg(k)} x[e,:] * u[e,c]
//   k_seg_mulbg:   out[k,c,:] = sum_{e in seg(k)} x[e,:] * rows3[idx[e],c,:]
//   k_ew_g3<0>:    out[e,:]   = sum_c u[e,c] * rows3[idx[e],c,:]                      (contract_c, gathered)
//   k_ew_g3<1>:    out[e,c]   = < rows3[idx[e],c,:], x[e,:] >                         (row_dot, gathered)
//   k_ew_g3<2>:    out[e,:]   = sum_c rows3[idx[e],c,:] * rowsb[idxb[e],c,:]          (sum_mul_c, both gathered)
template <bool GATHER>
__global__ void __launch_bounds__(256)
k_seg_prod(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ rows3, const int* __restrict__ idx,
           const int* __restrict__ perm, const int* __restrict__ row_ptr, int n_rows, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n_rows) return;
    float4 acc[3] = {f4_zero(), f4_zero(), f4_zero()};
    const int r0 = row_ptr[k], r1 = row_ptr[k + 1];
    for (int t = r0; t < r1; ++t) {
        const int e = perm ? perm[t] : t;
        const float4 xv = ld4(x + (size_t)e * kF + 4 * lane);
        if (GATHER) {
            const size_t g = (size_t)idx[e] * 3 * kF + 4 * lane;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = f4_fma(xv, ld4(rows3 + g + c * kF), acc[c]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = f4_fma(u[3 * e + c], xv, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) st4(out + ((size_t)k * 3 + c) * kF + 4 * lane, acc[c]);
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_ew_g3(const float* __restrict__ rows3, const int* __restrict__ idx, const float* __restrict__ p, const float* __restrict__ rowsb,
        const int* __restrict__ idxb, float* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const size_t g = (size_t)idx[e] * 3 * kF + 4 * lane, r = (size_t)e * kF + 4 * lane;
    if (MODE == 0) {            // p = u [n,3]
        float4 acc = f4_zero();
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = f4_fma(p[3 * e + c], ld4(rows3 + g + c * kF), acc);
        st4(out + r, acc);
    } else if (MODE == 1) {     // p = x [n,F]
        const float4 x = ld4(p + r);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = warp_sum(f4_dot(ld4(rows3 + g + c * kF), x));
            if (lane == 0) out[3 * e + c] = s;
        }
    } else {
        const size_t gb = (size_t)idxb[e] * 3 * kF + 4 * lane;
        float4 acc = f4_zero();
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = f4_fma(ld4(rows3 + g + c * kF), ld4(rowsb + gb + c * kF), acc);
        st4(out + r, acc);
    }
}

}  // namespace

extern "C" int nn_seg_prod(const float* x, const float* u, const float* rows3, const int32_t* idx, const int32_t* perm,
                           const int32_t* row_ptr, int32_t n_rows, float* out, void* stream) {
    NN_REQUIRE(x && row_ptr && out, "null pointer");
    NN_REQUIRE((u != nullptr) != (rows3 != nullptr), "exactly one of u (outer) and rows3 (gathered product)");
    NN_REQUIRE(!rows3 || idx, "rows3 needs idx");
    if (n_rows <= 0) return 0;
    const int grid = nn_ceil_div(n_rows, 8);
    if (rows3) k_seg_prod<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, u, rows3, idx, perm, row_ptr, n_rows, out);
    else k_seg_prod<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, u, rows3, idx, perm, row_ptr, n_rows, out);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_seg_prod");
    return 0;
}

extern "C" int nn_ew_g3(int32_t mode, const float* rows3, const int32_t* idx, const float* p, const float* rowsb, const int32_t* idxb,
                        float* out, int32_t n_rows, void* stream) {
    NN_REQUIRE(rows3 && idx && out, "null pointer");
    NN_REQUIRE(mode >= 0 && mode <= 2, "mode 0..2");
    NN_REQUIRE(mode == 2 ? (rowsb && idxb) : (p != nullptr), "missing operand");
    if (n_rows <= 0) return 0;
    const int grid = nn_ceil_div(n_rows, 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 0) k_ew_g3<0><<<grid, 256, 0, s>>>(rows3, idx, p, rowsb, idxb, out, n_rows);
    else if (mode == 1) k_ew_g3<1><<<grid, 256, 0, s>>>(rows3, idx, p, rowsb, idxb, out, n_rows);
    else k_ew_g3<2><<<grid, 256, 0, s>>>(rows3, idx, p, rowsb, idxb, out, n_rows);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_g3");
    return 0;
}

extern "C" int nn_ew_gmul(const float* a, const float* b, const float* r1, const int32_t* i1, const float* r2, const int32_t* i2,
                          float* out, int32_t n_rows, void* stream) {
    NN_REQUIRE(a && r1 && i1 && out, "null pointer");
    NN_REQUIRE((r2 == nullptr) == (i2 == nullptr), "r2 and i2 come together");
    if (n_rows <= 0) return 0;
    k_ew_gmul<<<nn_ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(a, b, r1, i1, r2, i2, out, n_rows); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_gmul");
    return 0;
}

extern "C" int nn_ew_mul3(const float* a, const float* b, const float* c, float* out, int64_t n_floats, void* stream) {
    NN_REQUIRE(a && b && c && out, "null pointer");
    NN_REQUIRE(n_floats % 4 == 0, "length must be a multiple of 4");
    if (n_floats <= 0) return 0;
    const long long n4 = n_floats / 4;
    long long g = (n4 + 255) / 256; if (g > nn_num_sms() * 16) g = nn_num_sms() * 16;
    k_ew_mul3<<<(int)g, 256, 0, (cudaStream_t)stream>>>(a, b, c, out, n4); NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_mul3");
    return 0;
}

extern "C" int nn_ew_rows(int32_t mode, const float* p, const float* q3, const float* u, float* out, int32_t n_rows, void* stream) {
    NN_REQUIRE(out != nullptr, "null pointer");
    NN_REQUIRE(mode >= 0 && mode <= 4, "mode 0..4");
    if (n_rows <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = nn_ceil_div(n_rows, 8);
    switch (mode) {
    case 0: NN_REQUIRE(p && u, "outer needs p, u"); k_ew_rows<0><<<grid, 256, 0, s>>>(p, q3, u, out, n_rows); break;
    case 1: NN_REQUIRE(q3 && u, "contract_c needs q3, u"); k_ew_rows<1><<<grid, 256, 0, s>>>(p, q3, u, out, n_rows); break;
    case 2: NN_REQUIRE(q3 && p, "row_dot needs q3, p"); k_ew_rows<2><<<grid, 256, 0, s>>>(p, q3, u, out, n_rows); break;
    case 3: NN_REQUIRE(q3 && p, "mul_b needs p, q3"); k_ew_rows<3><<<grid, 256, 0, s>>>(p, q3, u, out, n_rows); break;
    default: NN_REQUIRE(q3 && p, "sum_mul_c needs two [n,3,F] operands"); k_ew_rows<4><<<grid, 256, 0, s>>>(p, q3, u, out, n_rows); break;
    }
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_rows");
    return 0;
}

// ---------------------------------------------------------------------------- SiLU family (closed up to the double backward)
//   mode 0: out = silu(x)          mode 1: out = a * silu'(x)          mode 2: out = a * b * silu''(x)
// silu' = s (1 + x (1 - s)),  silu'' = s (1 - s) (2 + x (1 - 2 s)),  s = sigmoid(x).  d silu(x) = silu'(x) dx;
// d [a silu'(x)] = silu'(x) da + a silu''(x) dx: the double backward of an activation is two launches instead of the ~10
// element-wise ATen kernels autograd composes for it (23 activation sites per training step).
namespace {
template <int MODE>
__global__ void k_ew_silu(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                          float* __restrict__ out, long long n4) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const float4 xv = ld4(x + 4 * t);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float av[4] = {1.f, 1.f, 1.f, 1.f}, bv[4] = {1.f, 1.f, 1.f, 1.f}, o[4];
        if (MODE >= 1) { const float4 t4 = ld4(a + 4 * t); av[0] = t4.x; av[1] = t4.y; av[2] = t4.z; av[3] = t4.w; }
        if (MODE == 2) { const float4 t4 = ld4(b + 4 * t); bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float s = 1.0f / (1.0f + expf(-xs[k]));
            if (MODE == 0) o[k] = xs[k] * s;
            else if (MODE == 1) o[k] = av[k] * (s * fmaf(xs[k], 1.0f - s, 1.0f));
            else o[k] = av[k] * bv[k] * (s * (1.0f - s) * fmaf(xs[k], 1.0f - 2.0f * s, 2.0f));
        }
        st4(out + 4 * t, make_float4(o[0], o[1], o[2], o[3]));
    }
}
}  // namespace

extern "C" int nn_ew_silu(int32_t mode, const float* x, const float* a, const float* b, float* out, int64_t n_floats, void* stream) {
    NN_REQUIRE(x && out && (mode < 1 || a) && (mode < 2 || b), "null pointer");
    NN_REQUIRE(mode >= 0 && mode <= 2, "mode 0..2");
    NN_REQUIRE(n_floats % 4 == 0, "length must be a multiple of 4");
    if (n_floats <= 0) return 0;
    const long long n4 = n_floats / 4;
    long long g = (n4 + 255) / 256; if (g > nn_num_sms() * 16) g = nn_num_sms() * 16;
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 0) k_ew_silu<0><<<(int)g, 256, 0, s>>>(x, a, b, out, n4);
    else if (mode == 1) k_ew_silu<1><<<(int)g, 256, 0, s>>>(x, a, b, out, n4);
    else k_ew_silu<2><<<(int)g, 256, 0, s>>>(x, a, b, out, n4);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_silu");
    return 0;
}

// ---------------------------------------------------------------------------- radial basis with its first two x-derivatives
// R^k_n(x) = d^k/dx^k [ env(x) sin(f_n x) / x ],  env(x) = (1-x)^3 p(x) = 1 - 55 x^9 + 99 x^10 - 45 x^11  (PolynomialCutoff *
// RadialBesselLayer, layers/representations.py:166-169,233).  Two kernels closed under differentiation:
//   rbf_scale(k): out[e,n] = s[e] * R^k_n(x[e])          d/ds = rbf_dot(k)(g),  d/dx = s * rbf_dot(k+1)(g)
//   rbf_dot(k)  : out[e]   = sum_n g[e,n] R^k_n(x[e])    d/dg = rbf_scale(k)(go), d/dx = go * rbf_dot(k+1)(g)
// k = 0, 1, 2 (forward, forces, double backward).  At x = 1 (padding rows of the static edge list) env, env' and env''
// vanish exactly in the factored forms, so padded rows contribute exact zeros at every order.
namespace {
__device__ __forceinline__ void rbf_terms(float x, int k, float& e0, float& e1, float& e2) {
    float p = 45.f;
    p = fmaf(p, x, 36.f); p = fmaf(p, x, 28.f); p = fmaf(p, x, 21.f); p = fmaf(p, x, 15.f);
    p = fmaf(p, x, 10.f); p = fmaf(p, x, 6.f); p = fmaf(p, x, 3.f); p = fmaf(p, x, 1.f);
    const float t = 1.0f - x, x2 = x * x, x4 = x2 * x2, x7 = x4 * x2 * x;
    e0 = t * t * t * p;
    e1 = -495.f * x7 * x * t * t;
    e2 = -495.f * x7 * t * fmaf(-10.f, x, 8.f);
    (void)k;
}
// value of R^k_n at x for frequency f
__device__ __forceinline__ float rbf_k(float x, float f, int k, float e0, float e1, float e2) {
    float s, c;
    sincosf(f * x, &s, &c);
    const float ix = 1.0f / x;
    const float sb = s * ix;
    if (k == 0) return e0 * sb;
    const float sb1 = (f * x * c - s) * ix * ix;
    if (k == 1) return fmaf(e1, sb, e0 * sb1);
    const float sb2 = (-f * f * x * x * s - 2.f * f * x * c + 2.f * s) * ix * ix * ix;
    return fmaf(e2, sb, fmaf(2.f * e1, sb1, e0 * sb2));
}
__global__ void k_rbf_scale(const float* __restrict__ sc, const float* __restrict__ x, const float* __restrict__ freq, int k,
                            float* __restrict__ out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * kNB) return;
    const int e = t / kNB, j = t - e * kNB;
    const float xv = x[e];
    float e0, e1, e2;
    rbf_terms(xv, k, e0, e1, e2);
    out[t] = (sc ? sc[e] : 1.0f) * rbf_k(xv, freq[j], k, e0, e1, e2);
}
__global__ void k_rbf_dot(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ freq, int k,
                          float* __restrict__ out, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float xv = x[e];
    float e0, e1, e2;
    rbf_terms(xv, k, e0, e1, e2);
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < kNB; ++j) acc = fmaf(g[(size_t)e * kNB + j], rbf_k(xv, freq[j], k, e0, e1, e2), acc);
    out[e] = acc;
}
}  // namespace

extern "C" int nn_ew_rbf(int32_t op, int32_t k, const float* a, const float* x, const float* freq, float* out, int32_t n_rows, void* stream) {
    NN_REQUIRE(x && freq && out, "null pointer");
    NN_REQUIRE(k >= 0 && k <= 2, "derivative order 0..2");
    NN_REQUIRE(op == 0 || (op == 1 && a), "op 0 = scale (a optional), 1 = dot (a = g required)");
    if (n_rows <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (op == 0) k_rbf_scale<<<nn_ceil_div((long long)n_rows * kNB, 256), 256, 0, s>>>(a, x, freq, k, out, n_rows);
    else k_rbf_dot<<<nn_ceil_div(n_rows, 128), 128, 0, s>>>(a, x, freq, k, out, n_rows);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("nn_ew_rbf");
    return 0;
}
