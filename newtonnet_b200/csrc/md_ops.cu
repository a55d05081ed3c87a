// Device-resident molecular dynamics around the force path - SURVEY.md section 8f rank 1.
//
// The reference drives MD from ASE on the host (scripts/simulate.py:21-31: Langevin(atoms, 0.5 fs, 300 K) with
// MLAseCalculator.calculate, utils/ase_interface.py:52-81, called once per step: numpy -> torch -> H2D, forward,
// D2H -> numpy).  Here the integrator state (unwrapped positions, velocities: fp64) lives in HBM next to the
// force buffers; one step = nn_md_advance -> nn_nbr_count/fill -> nn_eval -> nn_md_finish, captured once as a
// CUDA graph and replayed without any host round trip.  The splitting is BAOAB (B half kick, A half drift,
// O exact Ornstein-Uhlenbeck, A half drift, force, B half kick); with ou_c = 1 it is velocity Verlet.
// Random numbers: Philox4x32-10 keyed by the seed, counter = (atom, step) - replaying a step reproduces it.
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// three standard normals for (atom, step): Box-Muller on two Philox blocks' worth of uniforms
__device__ __forceinline__ void normal3(unsigned long long seed, int atom, long long step, double (&g)[3]) {
    uint32_t c[4] = {(uint32_t)atom, (uint32_t)step, (uint32_t)((unsigned long long)step >> 32), 0x4D44u};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double two32 = 1.0 / 4294967296.0;
    const double u0 = ((double)c[0] + 0.5) * two32, u1 = ((double)c[1] + 0.5) * two32;
    const double u2 = ((double)c[2] + 0.5) * two32, u3 = ((double)c[3] + 0.5) * two32;
    const double r0 = sqrt(-2.0 * log(u0)), r1 = sqrt(-2.0 * log(u2));
    double s0, c0, s1, c1;
    sincospi(2.0 * u1, &s0, &c0);
    sincospi(2.0 * u3, &s1, &c1);
    g[0] = r0 * c0; g[1] = r0 * s0; g[2] = r1 * c1;
    (void)s1;
}

// B A O A + wrap.  x, v: [N,3] fp64 state; force [N,3] fp32 at the current x; pos_model: wrapped fp32 copy the
// neighbour search and the network read (reference: atoms.get_positions(wrap=True), ase_interface.py:135).
__global__ void k_md_advance(int n, double* __restrict__ x, double* __restrict__ v, const float* __restrict__ force,
                             const double* __restrict__ inv_mass, const float* __restrict__ cell,
                             const long long* __restrict__ batch, float* __restrict__ pos_model, double dt, double ou_c,
                             double kT, unsigned long long seed, const long long* __restrict__ step_ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double im = inv_mass[i], h = 0.5 * dt;
    double p[3], u[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        u[a] = v[3 * i + a] + h * im * (double)force[3 * i + a];
        p[a] = x[3 * i + a] + h * u[a];
    }
    if (ou_c < 1.0) {
        double g[3];
        normal3(seed, i, *step_ctr, g);
        const double sigma = sqrt((1.0 - ou_c * ou_c) * kT * im);
#pragma unroll
        for (int a = 0; a < 3; ++a) u[a] = ou_c * u[a] + sigma * g[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        p[a] += h * u[a];
        x[3 * i + a] = p[a];
        v[3 * i + a] = u[a];
    }
    // wrap into the cell: rows of `cell` are the lattice vectors, pos = frac @ cell
    const float* Cf = cell + 9 * batch[i];
    double c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = (double)Cf[k];
    const double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
    if (det != 0.0) {
        const double id = 1.0 / det;
        // inverse of cell (row-vector convention): frac = pos @ inv
        const double inv[9] = {(c[4] * c[8] - c[5] * c[7]) * id, (c[2] * c[7] - c[1] * c[8]) * id, (c[1] * c[5] - c[2] * c[4]) * id,
                               (c[5] * c[6] - c[3] * c[8]) * id, (c[0] * c[8] - c[2] * c[6]) * id, (c[2] * c[3] - c[0] * c[5]) * id,
                               (c[3] * c[7] - c[4] * c[6]) * id, (c[1] * c[6] - c[0] * c[7]) * id, (c[0] * c[4] - c[1] * c[3]) * id};
        double f[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            f[a] = p[0] * inv[a] + p[1] * inv[3 + a] + p[2] * inv[6 + a];
            f[a] -= floor(f[a]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = f[0] * c[a] + f[1] * c[3 + a] + f[2] * c[6 + a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) pos_model[3 * i + a] = (float)p[a];
}

// B half kick with the new force + per-system log row {potential, kinetic} + sticky status + step counter.
// One block per system, fixed-order tree reduction (deterministic).
__global__ void __launch_bounds__(256)
k_md_finish(const int* __restrict__ sys_ptr, double* __restrict__ v, const float* __restrict__ force,
            const double* __restrict__ inv_mass, double dt, const float* __restrict__ energy, double* __restrict__ log,
            int log_cap, long long* __restrict__ step_ctr, const int* __restrict__ nbr_status, int* __restrict__ sticky,
            unsigned int* __restrict__ ticket) {
    __shared__ double red[256];
    const int b = blockIdx.x, n_sys = gridDim.x;
    const long long step = *step_ctr;
    const double h = 0.5 * dt;
    double ke = 0.0;
    for (int i = sys_ptr[b] + threadIdx.x; i < sys_ptr[b + 1]; i += blockDim.x) {
        const double im = inv_mass[i];
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double u = v[3 * i + a] + h * im * (double)force[3 * i + a];
            v[3 * i + a] = u;
            s += u * u;
        }
        ke += im > 0.0 ? 0.5 * s / im : 0.0;
    }
    red[threadIdx.x] = ke;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double* row = log + ((size_t)(step % log_cap) * n_sys + b) * 2;
        row[0] = (double)energy[b];
        row[1] = red[0];
        if (b == 0) {
            if (nbr_status[NN_ST_EDGE_OVERFLOW]) atomicMax(&sticky[0], nbr_status[NN_ST_EDGE_OVERFLOW]);
            if (nbr_status[NN_ST_ROW_OVERFLOW]) atomicMax(&sticky[1], nbr_status[NN_ST_ROW_OVERFLOW]);
            if (nbr_status[NN_ST_SINGULAR_CELL]) sticky[2] = 1;
        }
        __threadfence();
        if (atomicAdd(ticket, 1u) == (unsigned)n_sys - 1) {     // last block: everyone has read the counter
            *ticket = 0;
            *step_ctr = step + 1;
        }
    }
}

}  // namespace

extern "C" {

NN_API int nn_md_advance(int32_t n_atoms, double* x, double* v, const float* force, const double* inv_mass, const float* cell,
                         const int64_t* batch, float* pos_model, double dt, double ou_c, double kT, uint64_t seed,
                         const int64_t* step_ctr, void* stream) {
    NN_REQUIRE(n_atoms >= 0 && x && v && force && inv_mass && cell && batch && pos_model && step_ctr, "null argument");
    NN_REQUIRE(ou_c >= 0.0 && ou_c <= 1.0 && kT >= 0.0, "ou_c must be in [0, 1] and kT >= 0");
    if (n_atoms == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    k_md_advance<<<(n_atoms + 127) / 128, 128, 0, s>>>(n_atoms, x, v, force, inv_mass, cell, (const long long*)batch, pos_model,
                                                       dt, ou_c, kT, (unsigned long long)seed, (const long long*)step_ctr);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("k_md_advance");
    return 0;
}

NN_API int nn_md_finish(int32_t n_systems, const int32_t* sys_ptr, double* v, const float* force, const double* inv_mass, double dt,
                        const float* energy, double* log, int32_t log_cap, int64_t* step_ctr, const int32_t* nbr_status,
                        int32_t* sticky, uint32_t* ticket, void* stream) {
    NN_REQUIRE(n_systems > 0 && sys_ptr && v && force && inv_mass && energy && log && log_cap > 0 && step_ctr && nbr_status &&
               sticky && ticket, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    k_md_finish<<<n_systems, 256, 0, s>>>(sys_ptr, v, force, inv_mass, dt, energy, log, log_cap, (long long*)step_ctr,
                                          nbr_status, sticky, ticket);
    NN_LAUNCHED(1);
    NN_CHECK_LAUNCH("k_md_finish");
    return 0;
}

}  // extern "C"
