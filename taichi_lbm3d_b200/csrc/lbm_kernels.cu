// Fused D3Q19 MRT kernels for sm_100a.
//
// One launch = one reference step() (Single_phase/LBM_3D_SinglePhase_Solver.py:477-481)
// for a range of nodes, in the order  pull-stream (:259-268, as a pull) -> face BCs
// (:272-370) -> macroscopic (:372-392) -> collision (:222-241) -> store.  The pipeline
// state kept in HBM is the post-collision populations f* (SoA, one plane per direction),
// double buffered (A-B): 19 fp32 read + 19 fp32 written per fluid node = 152 B.
//
//   MODE_STEP     the fused step
//   MODE_EXTRACT  pull + BC + macro only, written to the user-visible rho / v / F arrays
//                 (what to_numpy() shows after the reference's streaming3)
//   MODE_COLLIDE  collision of the user-visible state (first step after init or after a
//                 from_numpy): reads F, rho, v exactly as the reference's colission does
//
// Compiled twice (see lbm_kernels.cuh): namespace lbm_fast / lbm_strict.
#include "lbm_kernels.cuh"

#ifdef LBM_STRICT
#define LBM_NS lbm_strict
#else
#define LBM_NS lbm_fast
#endif

namespace LBM_NS {
using namespace d3q19;

__device__ __forceinline__ uint32_t vbc_slot(const StepArgs &a, int face, uint32_t lin) {
    const uint32_t z = lin % (uint32_t)a.nz;
    const uint32_t t = lin / (uint32_t)a.nz;
    const uint32_t y = t % (uint32_t)a.ny;
    const uint32_t x = t / (uint32_t)a.ny;
    const uint32_t s = face < 2 ? y * a.nz + z : (face < 4 ? x * a.nz + z : x * a.ny + y);
    return a.vbc_off[face] + s;
}

// Face BCs -> macro -> (collide).  `f` holds the streamed populations on entry.
template <bool FORCE, int MODE>
__device__ __forceinline__ void node_update(float (&f)[19], const StepArgs &a, uint32_t fl,
                                            uint32_t lin, float &rho, float &ux, float &uy,
                                            float &uz) {
    uint32_t slot = 0;
    bool pressure = false;
    if (a.has_bc) {
        const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
        if (bc) {
            const int face = (int)bc - 1;
            const int type = a.P.bc_type[face];
            if (type == 1) {                      // :274-281  F = feq(rho_bc, v_prev)
                float u0 = 0.f, u1 = 0.f, u2 = 0.f;
                slot = vbc_slot(a, face, lin);
                if (!(fl & FL_PIN_SOLID)) {
                    u0 = a.vbc[3 * (size_t)slot + 0];
                    u1 = a.vbc[3 * (size_t)slot + 1];
                    u2 = a.vbc[3 * (size_t)slot + 2];
                }
                feq_all(f, a.P.bc_rho[face], u0, u1, u2);
                pressure = true;
            } else if (type == 2) {               // :283-288  F = feq(1, bc_vel)
                feq_all(f, 1.0f, a.P.bc_vel[face][0], a.P.bc_vel[face][1], a.P.bc_vel[face][2]);
            }
        }
    }
    macro(f, a.P, FORCE, rho, ux, uy, uz);
    if (MODE == MODE_STEP) {
        if (pressure) {
            a.vbc[3 * (size_t)slot + 0] = ux;
            a.vbc[3 * (size_t)slot + 1] = uy;
            a.vbc[3 * (size_t)slot + 2] = uz;
        }
        collide(f, a.P, FORCE, rho, ux, uy, uz);
    }
}

// MODE_COLLIDE: the reference's colission() on the user-visible state.
template <bool FORCE>
__device__ __forceinline__ void node_collide_only(float (&f)[19], const StepArgs &a, uint32_t fl,
                                                  uint32_t lin) {
    float rho = 1.0f, ux = 0.f, uy = 0.f, uz = 0.f;
    if (a.F != nullptr) {
#pragma unroll
        for (int s = 0; s < 19; ++s) f[s] = a.F[(size_t)lin * 19 + s];
        rho = a.rho[lin];
        ux = a.v[(size_t)lin * 3 + 0];
        uy = a.v[(size_t)lin * 3 + 1];
        uz = a.v[(size_t)lin * 3 + 2];
    } else {                                      // pristine init(): F = w, rho = 1, v = 0
#pragma unroll
        for (int s = 0; s < 19; ++s) f[s] = weight(s);
    }
    if (a.has_bc) {                               // seed v_prev of pressure-BC nodes
        const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
        if (bc && a.P.bc_type[bc - 1] == 1) {
            const uint32_t slot = vbc_slot(a, (int)bc - 1, lin);
            a.vbc[3 * (size_t)slot + 0] = ux;
            a.vbc[3 * (size_t)slot + 1] = uy;
            a.vbc[3 * (size_t)slot + 2] = uz;
        }
    }
    collide(f, a.P, FORCE, rho, ux, uy, uz);
}

template <int MODE>
__device__ __forceinline__ void write_user_fields(const StepArgs &a, uint32_t lin,
                                                  const float (&f)[19], float rho, float ux,
                                                  float uy, float uz) {
    a.rho[lin] = rho;
    a.v[(size_t)lin * 3 + 0] = ux;
    a.v[(size_t)lin * 3 + 1] = uy;
    a.v[(size_t)lin * 3 + 2] = uz;
    if (a.F != nullptr) {
#pragma unroll
        for (int s = 0; s < 19; ++s) a.F[(size_t)lin * 19 + s] = f[s];
    }
}

// ---------------------------------------------------------------------------------------------
// dense lattice: direct addressing, node = linear index i*ny*nz + j*nz + k (z fastest)
// ---------------------------------------------------------------------------------------------
template <bool FORCE, int MODE>
__global__ void __launch_bounds__(256) k_dense(const StepArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const uint32_t idx = a.first + t;
    const uint32_t fl = a.flags[idx];
    if (fl & FL_SOLID) return;        // solid nodes keep f = F = w, rho = 1, v = 0 (:164-169, :390-392)
    const size_t N = a.stride;
    float f[19];
    if (MODE == MODE_COLLIDE) {
        node_collide_only<FORCE>(f, a, fl, idx);
    } else {
        const float *__restrict__ p = a.fin + idx;
        const int sx = a.ny * a.nz, sy = a.nz;
        // offsets to the x-1 / x+1 ... neighbours with the periodic wrap of :247-257
        const int oxm = (fl & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
        const int oxp = (fl & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
        const int oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
        const int oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
        const int ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
        const int ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
        // pull: F[i][s] = f*[i - e_s][s]; source offset per direction
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
        if ((fl & FL_LINK_MASK) == 0) {
#define X(s, ex, ey, ez, o) f[s] = __ldg(p + (size_t)s * N + OFF(ex, ey, ez));
            D3Q19_DIRS(X)
#undef X
        } else {
            // half-way bounce-back (:267-268): source solid -> own opposite population
#define X(s, ex, ey, ez, o)                                                                    \
    f[s] = ((fl >> s) & 1u) ? __ldg(p + (size_t)o * N) : __ldg(p + (size_t)s * N + OFF(ex, ey, ez));
            D3Q19_DIRS(X)
#undef X
        }
#undef OFF
        float rho, ux, uy, uz;
        node_update<FORCE, MODE>(f, a, fl, idx, rho, ux, uy, uz);
        if (MODE == MODE_EXTRACT) {
            write_user_fields<MODE>(a, idx, f, rho, ux, uy, uz);
            return;
        }
    }
    float *__restrict__ q = a.fout + idx;
#pragma unroll
    for (int s = 0; s < 19; ++s) q[(size_t)s * N] = f[s];
}

// ---------------------------------------------------------------------------------------------
// sparse storage: compacted fluid list + 18-neighbour pull table
// ---------------------------------------------------------------------------------------------
template <bool FORCE, int MODE>
__global__ void __launch_bounds__(256) k_sparse(const StepArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const uint32_t i = a.first + t;
    const size_t N = a.stride;
    const uint32_t fl = a.has_bc ? a.flags[i] : 0u;
    float f[19];
    if (MODE == MODE_COLLIDE) {
        node_collide_only<FORCE>(f, a, fl, a.lin[i]);
    } else {
        const float *__restrict__ p = a.fin;
        const int32_t *__restrict__ nb = a.nbr + i;
        f[0] = __ldg(p + i);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        const int32_t j = __ldg(nb + (size_t)(s - 1) * N);                                     \
        f[s] = j >= 0 ? __ldg(p + (size_t)s * N + j) : __ldg(p + (size_t)o * N + i);           \
    }
        D3Q19_DIRS(X)
#undef X
        float rho, ux, uy, uz;
        const uint32_t lin = (MODE == MODE_EXTRACT || a.has_bc) ? a.lin[i] : 0u;
        node_update<FORCE, MODE>(f, a, fl, lin, rho, ux, uy, uz);
        if (MODE == MODE_EXTRACT) {
            write_user_fields<MODE>(a, lin, f, rho, ux, uy, uz);
            return;
        }
    }
    float *__restrict__ q = a.fout + i;
#pragma unroll
    for (int s = 0; s < 19; ++s) q[(size_t)s * N] = f[s];
}

#define LAUNCH_CASE(KERN, F, M)                                                                \
    KERN<F, M><<<grid, block, 0, st>>>(a);                                                     \
    break;

static cudaError_t launch_any(bool sparse, int mode, const StepArgs &a, int block, cudaStream_t st) {
    if (a.count == 0) return cudaSuccess;
    if (block <= 0 || block > 256) block = 256;
    const unsigned grid = (a.count + block - 1) / block;
    const int key = (sparse ? 8 : 0) | (a.force ? 4 : 0) | mode;
    switch (key) {
        case 0: LAUNCH_CASE(k_dense, false, MODE_STEP)
        case 1: LAUNCH_CASE(k_dense, false, MODE_EXTRACT)
        case 2: LAUNCH_CASE(k_dense, false, MODE_COLLIDE)
        case 4: LAUNCH_CASE(k_dense, true, MODE_STEP)
        case 5: LAUNCH_CASE(k_dense, true, MODE_EXTRACT)
        case 6: LAUNCH_CASE(k_dense, true, MODE_COLLIDE)
        case 8: LAUNCH_CASE(k_sparse, false, MODE_STEP)
        case 9: LAUNCH_CASE(k_sparse, false, MODE_EXTRACT)
        case 10: LAUNCH_CASE(k_sparse, false, MODE_COLLIDE)
        case 12: LAUNCH_CASE(k_sparse, true, MODE_STEP)
        case 13: LAUNCH_CASE(k_sparse, true, MODE_EXTRACT)
        case 14: LAUNCH_CASE(k_sparse, true, MODE_COLLIDE)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_dense(int mode, const StepArgs &a, int block, cudaStream_t st) {
    return launch_any(false, mode, a, block, st);
}
cudaError_t launch_sparse(int mode, const StepArgs &a, int block, cudaStream_t st) {
    return launch_any(true, mode, a, block, st);
}

cudaError_t set_inverse_matrix(const float *invM361) {
#ifdef LBM_STRICT
    return cudaMemcpyToSymbol(c_invM, invM361, 361 * sizeof(float));
#else
    (void)invM361;
    return cudaSuccess;
#endif
}

}  // namespace LBM_NS
