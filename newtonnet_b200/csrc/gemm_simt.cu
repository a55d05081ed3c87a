// fp32 SIMT contraction Y[M,128] = epi(pro(X)[M,128] @ B[128,128]) - the exact-fp32 baseline backend of
// nn_gemm128 (backend 0).  The tensor-core backend (gemm_tc.cu, tcgen05 3xTF32) implements the same
// prologue / epilogue contract and is checked against this one.
//
// Replaces the cuBLAS SGEMM calls behind nn.Linear in models/newtonnet.py:209,218,222,230 and
// models/output.py:98-100 (forward) and their autograd transposes (backward).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, KTOT = 128;
constexpr int APAD = 4;

struct RowCount {
    const int* m_dev; int mul; int m;
    __device__ int get() const {
        if (m_dev == nullptr) return m;
        long long v = (long long)m_dev[0] * mul;
        return v < m ? (int)v : m;
    }
};

template <int PRO>
__device__ __forceinline__ float4 load_a(const nn_gemm_args& a, int grow, int k, int M) {
    if (grow >= M) return f4_zero();
    float4 v = ld4(a.X + (size_t)grow * KTOT + k);
    if (PRO == NN_PRO_SILU) {
        v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
    } else if (PRO == NN_PRO_SILU_SAVE) {   // every X element is loaded by exactly one thread of one block
        st4(a.aux_out + (size_t)grow * KTOT + k, make_float4(dsilu_f(v.x), dsilu_f(v.y), dsilu_f(v.z), dsilu_f(v.w)));
        v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
    } else if (PRO == NN_PRO_ROWSCALE3) {
        float4 s = ld4(a.aux2 + (size_t)(grow / 3) * KTOT + k);
        v = f4_mul(v, s);
    }
    return v;
}

template <int EPI>
__device__ __forceinline__ float4 epilogue(const nn_gemm_args& a, float4 acc, int grow, int col) {
    if (EPI == NN_EPI_BIAS) {
        if (a.bias) acc = f4_add(acc, ld4(a.bias + col));
    } else if (EPI == NN_EPI_DSILU) {
        float4 p = ld4(a.aux1 + (size_t)grow * BN + col);
        acc.x *= dsilu_f(p.x); acc.y *= dsilu_f(p.y); acc.z *= dsilu_f(p.z); acc.w *= dsilu_f(p.w);
    } else if (EPI == NN_EPI_ADD) {
        acc = f4_add(acc, ld4(a.aux1 + (size_t)grow * BN + col));
    } else if (EPI == NN_EPI_MUL) {
        acc = f4_mul(acc, ld4(a.aux1 + (size_t)grow * BN + col));
    } else if (EPI == NN_EPI_EQUIV_BWD) {
        float4 fb = ld4(a.aux1 + (size_t)grow * BN + col);
        float4 ab = ld4(a.aux2 + (size_t)(grow / 3) * BN + col);
        float4 g = ld4(a.aux3 + (size_t)grow * BN + col);
        acc = f4_add(acc, f4_fma(ab, g, fb));
    }
    return acc;
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(256, 2) k_gemm128_simt(nn_gemm_args a) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];
    RowCount rc{a.m_dev, a.m_dev_mul, a.m};
    const int M = rc.get();
    const int row0 = blockIdx.x * BM;
    if (row0 >= M) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int idx = tid + r * 256;
            ra[r] = load_a<PRO>(a, row0 + (idx >> 2), k0 + (idx & 3) * 4, M);
            rb[r] = ld4(a.B + (size_t)(k0 + (idx >> 5)) * BN + (idx & 31) * 4);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int idx = tid + r * 256;
            int row = idx >> 2, kq = (idx & 3) * 4;
            As[buf][kq + 0][row] = ra[r].x; As[buf][kq + 1][row] = ra[r].y;
            As[buf][kq + 2][row] = ra[r].z; As[buf][kq + 3][row] = ra[r].w;
            st4(&Bs[buf][idx >> 5][(idx & 31) * 4], rb[r]);
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    constexpr int NKT = KTOT / BK;
#pragma unroll 1
    for (int kt = 0; kt < NKT; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < NKT) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = ld4(&As[cur][k][ty * 4]), a1 = ld4(&As[cur][k][64 + ty * 4]);
            float4 b0 = ld4(&Bs[cur][k][tx * 4]), b1 = ld4(&Bs[cur][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < NKT) {
            sstore(cur ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int grow = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (grow >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int col = h * 64 + tx * 4;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            v = epilogue<EPI>(a, v, grow, col);
            st4(a.Y + (size_t)grow * BN + col, v);
        }
    }
}

template <int PRO, int EPI>
void launch(const nn_gemm_args& a, cudaStream_t s) {
    k_gemm128_simt<PRO, EPI><<<nn_ceil_div(a.m, BM), 256, 0, s>>>(a); NN_LAUNCHED(1);
}

}  // namespace

int nn_gemm128_simt_launch(const nn_gemm_args& a, cudaStream_t s) {
    if (a.m <= 0) return 0;
#define NN_CASE(P, E) if (a.prologue == P && a.epilogue == E) { launch<P, E>(a, s); goto done; }
    NN_CASE(NN_PRO_NONE, NN_EPI_BIAS)
    NN_CASE(NN_PRO_SILU, NN_EPI_BIAS)
    NN_CASE(NN_PRO_NONE, NN_EPI_DSILU)
    NN_CASE(NN_PRO_NONE, NN_EPI_ADD)
    NN_CASE(NN_PRO_ROWSCALE3, NN_EPI_EQUIV_BWD)
    NN_CASE(NN_PRO_SILU_SAVE, NN_EPI_BIAS)
    NN_CASE(NN_PRO_NONE, NN_EPI_MUL)
#undef NN_CASE
    nn_set_error("nn_gemm128: unsupported prologue/epilogue combination %d/%d", a.prologue, a.epilogue);
    return -1;
done:
    NN_CHECK_LAUNCH("nn_gemm128(simt)");
    return 0;
}
