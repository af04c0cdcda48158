// Geometry preprocessing and small utility kernels shared by the single-phase and two-phase
// C-ABI layers (each translation unit gets its own copy: everything is in an anonymous
// namespace).  Reference: Single_phase/LBM_3D_SinglePhase_Solver.py (line numbers below).
#pragma once
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lbm_kernels.cuh"

namespace {

// ---- geometry preprocessing ------------------------------------------------------------------
struct GeoParams {
    int nx, ny, nz;
    int halo_x;
    int xface0, xface1;
    int bc_type[6];
    int two_phase;          // also flag nodes whose phase-field stencil touches a solid
    int vel_in_place;       // velocity faces use the scripts' in-place form (not an overwrite): the BC
                            // word records the last PRESSURE face, velocity faces go by the at-face bits
    int bc_psi_type[6];     // two-phase: 0 periodic / 1 constant psi per face (clamped stencil)
};

__constant__ int c_e[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
    {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1}, {1, 0, -1},
    {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};

// periodic_index :247-257 (x wrap disabled when ghost planes supply the neighbours)
__device__ __forceinline__ bool pull_source(const GeoParams &g, int x, int y, int z, int s, size_t &src) {
    int xs = x - c_e[s][0], ys = y - c_e[s][1], zs = z - c_e[s][2];
    if (g.halo_x) {
        if (xs < 0 || xs > g.nx - 1) return false;
    } else {
        if (xs < 0) xs = g.nx - 1;
        if (xs > g.nx - 1) xs = 0;
    }
    if (ys < 0) ys = g.ny - 1;
    if (ys > g.ny - 1) ys = 0;
    if (zs < 0) zs = g.nz - 1;
    if (zs > g.nz - 1) zs = 0;
    src = ((size_t)xs * g.ny + ys) * g.nz + zs;
    return true;
}

// BC bits of a fluid node: winning face (later face overwrites, :272-370) and whether a
// pressure face reads the zero velocity of a solid inward neighbour (:278, :294 ...).
__device__ __forceinline__ uint32_t bc_word(const GeoParams &g, const int8_t *solid, int x, int y, int z) {
    // single phase: both BC types overwrite all 19 populations, so only the last matching face
    // counts.  Two-phase: the velocity form (2phase/lbm_solver_3d_2phase.py:500-504) depends on
    // the current F, so the word records the last PRESSURE face and the kernel then applies
    // the later velocity faces in order from the at-face bits.
    int win = -1;
    const bool only_p = g.two_phase != 0 || g.vel_in_place != 0;
#define BC_MATCH(f) (g.bc_type[f] && (!only_p || g.bc_type[f] == 1))
    if (BC_MATCH(0) && x == g.xface0) win = 0;
    if (BC_MATCH(1) && x == g.xface1) win = 1;
    if (BC_MATCH(2) && y == 0) win = 2;
    if (BC_MATCH(3) && y == g.ny - 1) win = 3;
    if (BC_MATCH(4) && z == 0) win = 4;
    if (BC_MATCH(5) && z == g.nz - 1) win = 5;
#undef BC_MATCH
    if (win < 0) return 0u;
    uint32_t w = (uint32_t)(win + 1) << FL_BC_SHIFT;
    if (g.bc_type[win] == 1) {
        int xi = x, yi = y, zi = z;
        switch (win) {
            case 0: xi = x + 1; break;
            case 1: xi = x - 1; break;
            case 2: yi = 1; break;
            case 3: yi = g.ny - 2; break;
            case 4: zi = 1; break;
            default: zi = g.nz - 2; break;
        }
        const bool inside = xi >= 0 && xi < g.nx && yi >= 0 && yi < g.ny && zi >= 0 && zi < g.nz;
        if (inside && solid[((size_t)xi * g.ny + yi) * g.nz + zi] > 0) w |= FL_PIN_SOLID;
    }
    return w;
}

// bits 24..29: the node sits on a lattice face.  Without ghost planes these are the faces
// periodic_index wraps across (:247-257); in an x-slab of the two-phase solver the x bits mark the
// GLOBAL x faces (velocity and psi BCs, clamped psi stencil) -- the kernels never wrap x when
// ghost planes exist.
__device__ __forceinline__ uint32_t at_face_bits(const GeoParams &g, int x, int y, int z) {
    uint32_t fl = 0;
    if (!g.halo_x) {
        if (x == 0) fl |= FL_AT_X0;
        if (x == g.nx - 1) fl |= FL_AT_X1;
    } else if (g.two_phase || g.vel_in_place) {
        if (x == g.xface0) fl |= FL_AT_X0;
        if (x == g.xface1) fl |= FL_AT_X1;
    }
    if (y == 0) fl |= FL_AT_Y0;
    if (y == g.ny - 1) fl |= FL_AT_Y1;
    if (z == 0) fl |= FL_AT_Z0;
    if (z == g.nz - 1) fl |= FL_AT_Z1;
    return fl;
}

// two-phase: Compute_C (2phase/lbm_solver_3d_2phase.py:259-275) looks at i + e_s with
// periodic_index_for_psi (:390-428): wrap on periodic psi faces, clamp on constant ones
__device__ __forceinline__ uint32_t near_solid_bit(const GeoParams &g, const int8_t *solid, int x, int y, int z) {
    const int n[3] = {g.nx, g.ny, g.nz};
    for (int s = 1; s < 19; ++s) {
        int q[3] = {x + c_e[s][0], y + c_e[s][1], z + c_e[s][2]};
        for (int d = g.halo_x ? 1 : 0; d < 3; ++d) {
            if (q[d] < 0) q[d] = g.bc_psi_type[2 * d] == 0 ? n[d] - 1 : 0;
            if (q[d] > n[d] - 1) q[d] = g.bc_psi_type[2 * d + 1] == 0 ? 0 : n[d] - 1;
        }
        if (g.halo_x) {
            // ghost planes carry the periodic images; a constant-psi GLOBAL face clamps
            if (x == g.xface0 && q[0] < x && g.bc_psi_type[0] != 0) q[0] = x;
            if (x == g.xface1 && q[0] > x && g.bc_psi_type[1] != 0) q[0] = x;
            if (q[0] < 0 || q[0] > g.nx - 1) continue;      // ghost node itself: never updated
        }
        if (solid[((size_t)q[0] * g.ny + q[1]) * g.nz + q[2]] != 0) return FL_NEAR_SOLID;
    }
    return 0u;
}

__global__ void k_build_flags(const GeoParams g, const int8_t *__restrict__ solid,
                              uint32_t *__restrict__ flags, uint8_t *__restrict__ cls) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t N = (size_t)g.nx * g.ny * g.nz;
    if (idx >= N) return;
    const int z = (int)(idx % g.nz);
    const size_t t = idx / g.nz;
    const int y = (int)(t % g.ny), x = (int)(t / g.ny);
    uint32_t fl = 0;
    if (solid[idx] != 0) {
        flags[idx] = FL_SOLID;
        // a solid node in the same 32-byte sector (8 nodes of a z-row) as a fluid node stores
        // too, so the sector is written whole
        const int z0 = z & ~7;
        bool any_fluid = false;
        for (int q = z0; q < z0 + 8 && q < g.nz; ++q)
            if (solid[idx - z + q] == 0) any_fluid = true;
        cls[idx] = any_fluid ? NODE_SOLID_WRITE : NODE_SOLID;
        return;
    }
    for (int s = 1; s < 19; ++s) {
        size_t src;
        if (!pull_source(g, x, y, z, s, src) || solid[src] != 0) fl |= 1u << s;
    }
    fl |= at_face_bits(g, x, y, z);
    fl |= bc_word(g, solid, x, y, z);
    if (g.two_phase) fl |= near_solid_bit(g, solid, x, y, z);
    flags[idx] = fl;
    cls[idx] = fl == 0 ? NODE_BULK : NODE_SPECIAL;
}

struct IsFluid {
    __host__ __device__ uint32_t operator()(const int8_t &s) const { return s == 0 ? 1u : 0u; }
};

struct IsBulk {
    __host__ __device__ uint32_t operator()(const uint8_t &c) const { return c == NODE_BULK ? 1u : 0u; }
};

__global__ void k_fill(float *p, size_t n, float v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_fill_weights(float *F, size_t n_nodes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes * 19) F[i] = d3q19::weight((int)(i % 19));
}

__global__ void k_binarize(int8_t *s, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = s[i] > 0 ? 1 : 0;       // init_geo :175  in_dat[in_dat>0] = 1
}

// cal_max_v :399-402  (norm evaluated without FMA contraction so every mode agrees).
// out[0]: int image of the maximum (atomicMax on it orders non-negative floats), out[1]: set when
// any |v| is NaN -- fmaxf would drop it silently and a diverged run would report a finite speed.
__global__ void k_max_v(const float *__restrict__ v, size_t n, float *out) {
    float best = -1e10f;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
        const float nr = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        bad |= nr != nr;
        best = fmaxf(best, nr);
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (best >= 0.f) atomicMax((int *)out, __float_as_int(best));
        if (bad) atomicOr((int *)out + 1, 1);
    }
}

// host side of cal_max_v, shared by both solvers: reduction on `st`, result (NaN if any node is)
inline cudaError_t max_v_reduce(const float *d_v, size_t n, float *d_scalar2, cudaStream_t st, float *result) {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const int init[2] = {(int)0x80000000, 0};       // below every non-negative float; no NaN seen
    cudaError_t e = cudaMemcpyAsync(d_scalar2, init, sizeof init, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    k_max_v<<<n_sm * 8, 256, 0, st>>>(d_v, n, d_scalar2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    int got[2] = {0, 0};
    if ((e = cudaMemcpyAsync(got, d_scalar2, sizeof got, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    float v;
    memcpy(&v, &got[0], sizeof v);
    *result = got[1] ? nanf("") : (got[0] < 0 ? -1e10f : v);     // seed :395
    return cudaSuccess;
}

inline unsigned nblocks(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

// halo staging: 5 populations of one lattice plane <-> contiguous buffer [5][count]
struct HaloDirs { int s[5]; };
// generic over both storage modes: node i of the plane lives at  plane_s + (row0 + i/nz)*prow + i%nz
// (dense; SoA has prow = nz) or at  plane_s + first + i  (sparse: nz = 0)
__global__ void k_halo_pack(StepArgs a, uint32_t row0, uint32_t first, uint32_t count, HaloDirs d,
                            float *__restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t e = a.nz ? (row0 + i / (uint32_t)a.nz) * a.prow + i % (uint32_t)a.nz : first + i;
#pragma unroll
    for (int q = 0; q < 5; ++q) dst[(size_t)q * count + i] = a.pown[d.s[q]][e];
}
__global__ void k_halo_unpack(StepArgs a, uint32_t row0, uint32_t first, uint32_t count, HaloDirs d,
                              const float *__restrict__ src) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t e = a.nz ? (row0 + i / (uint32_t)a.nz) * a.prow + i % (uint32_t)a.nz : first + i;
#pragma unroll
    for (int q = 0; q < 5; ++q) a.pout[d.s[q]][e] = src[(size_t)q * count + i];
}

// both sides of a slab in ONE launch (blockIdx.y = side): the native slab loop packs the e_x = -1
// populations of the first owned plane and the e_x = +1 populations of the last owned plane
// together, and fills both ghost planes together
struct HaloSide { uint32_t row0, first, count; HaloDirs d; float *buf; };
__global__ void k_halo_pack2(StepArgs a, HaloSide s0, HaloSide s1) {
    const HaloSide &h = blockIdx.y == 0 ? s0 : s1;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.count) return;
    const uint32_t e = a.nz ? (h.row0 + i / (uint32_t)a.nz) * a.prow + i % (uint32_t)a.nz : h.first + i;
#pragma unroll
    for (int q = 0; q < 5; ++q) h.buf[(size_t)q * h.count + i] = a.pown[h.d.s[q]][e];
}
__global__ void k_halo_unpack2(StepArgs a, HaloSide s0, HaloSide s1) {
    const HaloSide &h = blockIdx.y == 0 ? s0 : s1;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.count) return;
    const uint32_t e = a.nz ? (h.row0 + i / (uint32_t)a.nz) * a.prow + i % (uint32_t)a.nz : h.first + i;
#pragma unroll
    for (int q = 0; q < 5; ++q) a.pout[h.d.s[q]][e] = h.buf[(size_t)q * h.count + i];
}

// exact inverse of M (:64-83) as rationals; every non-zero entry rounds to the same f32 as
// np.linalg.inv's (tests/test_abi_cpu.py); LAPACK's 1e-17 noise entries are exactly 0 here.
const double kInvM[19][19] = {
    {1.0/3.0, -1.0/2.0, 1.0/6.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, 1.0/6.0, -1.0/6.0, 0, 0, 0, 0, 1.0/12.0, -1.0/12.0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, -1.0/6.0, 1.0/6.0, 0, 0, 0, 0, 1.0/12.0, -1.0/12.0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, 0, 0, 1.0/6.0, -1.0/6.0, 0, 0, -1.0/24.0, 1.0/24.0, 1.0/8.0, -1.0/8.0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, 0, 0, -1.0/6.0, 1.0/6.0, 0, 0, -1.0/24.0, 1.0/24.0, 1.0/8.0, -1.0/8.0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, 0, 0, 0, 0, 1.0/6.0, -1.0/6.0, -1.0/24.0, 1.0/24.0, -1.0/8.0, 1.0/8.0, 0, 0, 0, 0, 0, 0},
    {1.0/18.0, 0, -1.0/18.0, 0, 0, 0, 0, -1.0/6.0, 1.0/6.0, -1.0/24.0, 1.0/24.0, -1.0/8.0, 1.0/8.0, 0, 0, 0, 0, 0, 0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 1.0/12.0, 1.0/24.0, 1.0/12.0, 1.0/24.0, 0, 0, 1.0/48.0, 1.0/48.0, 1.0/16.0, 1.0/16.0, 1.0/4.0, 0, 0, 1.0/8.0, -1.0/8.0, 0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, -1.0/12.0, -1.0/24.0, -1.0/12.0, -1.0/24.0, 0, 0, 1.0/48.0, 1.0/48.0, 1.0/16.0, 1.0/16.0, 1.0/4.0, 0, 0, -1.0/8.0, 1.0/8.0, 0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 1.0/12.0, 1.0/24.0, -1.0/12.0, -1.0/24.0, 0, 0, 1.0/48.0, 1.0/48.0, 1.0/16.0, 1.0/16.0, -1.0/4.0, 0, 0, 1.0/8.0, 1.0/8.0, 0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, -1.0/12.0, -1.0/24.0, 1.0/12.0, 1.0/24.0, 0, 0, 1.0/48.0, 1.0/48.0, 1.0/16.0, 1.0/16.0, -1.0/4.0, 0, 0, -1.0/8.0, -1.0/8.0, 0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 1.0/12.0, 1.0/24.0, 0, 0, 1.0/12.0, 1.0/24.0, 1.0/48.0, 1.0/48.0, -1.0/16.0, -1.0/16.0, 0, 0, 1.0/4.0, -1.0/8.0, 0, 1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, -1.0/12.0, -1.0/24.0, 0, 0, -1.0/12.0, -1.0/24.0, 1.0/48.0, 1.0/48.0, -1.0/16.0, -1.0/16.0, 0, 0, 1.0/4.0, 1.0/8.0, 0, -1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 1.0/12.0, 1.0/24.0, 0, 0, -1.0/12.0, -1.0/24.0, 1.0/48.0, 1.0/48.0, -1.0/16.0, -1.0/16.0, 0, 0, -1.0/4.0, -1.0/8.0, 0, -1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, -1.0/12.0, -1.0/24.0, 0, 0, 1.0/12.0, 1.0/24.0, 1.0/48.0, 1.0/48.0, -1.0/16.0, -1.0/16.0, 0, 0, -1.0/4.0, 1.0/8.0, 0, 1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 0, 0, 1.0/12.0, 1.0/24.0, 1.0/12.0, 1.0/24.0, -1.0/24.0, -1.0/24.0, 0, 0, 0, 1.0/4.0, 0, 0, 1.0/8.0, -1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 0, 0, -1.0/12.0, -1.0/24.0, -1.0/12.0, -1.0/24.0, -1.0/24.0, -1.0/24.0, 0, 0, 0, 1.0/4.0, 0, 0, -1.0/8.0, 1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 0, 0, 1.0/12.0, 1.0/24.0, -1.0/12.0, -1.0/24.0, -1.0/24.0, -1.0/24.0, 0, 0, 0, -1.0/4.0, 0, 0, 1.0/8.0, 1.0/8.0},
    {1.0/36.0, 1.0/24.0, 1.0/72.0, 0, 0, -1.0/12.0, -1.0/24.0, 1.0/12.0, 1.0/24.0, -1.0/24.0, -1.0/24.0, 0, 0, 0, -1.0/4.0, 0, 0, -1.0/8.0, -1.0/8.0}};

}  // namespace
