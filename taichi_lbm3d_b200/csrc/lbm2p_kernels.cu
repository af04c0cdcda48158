// Two-phase colour-gradient D3Q19 MRT kernels for sm_100a (dense storage).
//
// Reference: 2phase/lbm_solver_3d_2phase.py, main loop :626-632
//     colission :302-372 ; streaming1 :431 ; Boundary_condition :491 ; streaming3 :587 ;
//     Boundary_condition_psi :445
// Two launches per step instead of five passes:
//
//   k2p_colour   rho_r, rho_b of step n from the colour records of the 18 pull sources
//                (the recoloured g_r, g_b of :345-363 are RE-EVALUATED from the source node's
//                record (rho_r, rho_b, v, C) instead of being stored: 8 words per node instead
//                of 38, and a deterministic ascending-s sum instead of the reference's
//                unordered float atomics :365-372); then psi (:605) and the psi BC (:445-486).
//   k2p_main     pull-stream f* + flow BCs + macro of step n, then the collision of step n+1:
//                C = grad(psi) over 18 neighbours (:259-275), surface-tension perturbation of
//                meq (:316-321), psi-dependent relaxation (:278-299), Guo force (:241-247,
//                closed form), inv_M; writes f* and the node's colour record.
//
// The pull needs psi of step n at all neighbours before C can be formed, hence two kernels.
// Algorithmic bytes per node-step: 152 (populations) + 16 (rho_r, rho_b r/w) +
// 8 (psi r/w) = 176; this implementation moves 176 + 64 (record write + read) = 240.
#include <cstdlib>

#include "lbm2p_kernels.cuh"
#include "lbm_sparse_table.cuh"

#ifdef LBM_STRICT
#define LBM2P_NS lbm2p_strict
#else
#define LBM2P_NS lbm2p_fast
#endif

namespace LBM2P_NS {
using namespace d3q19;

__device__ __forceinline__ uint32_t vbc_slot2(const StepArgs &a, int face, uint32_t lin) {
    const uint32_t z = lin % (uint32_t)a.nz;
    const uint32_t t = lin / (uint32_t)a.nz;
    const uint32_t y = t % (uint32_t)a.ny;
    const uint32_t x = t / (uint32_t)a.ny;
    const uint32_t s = face < 2 ? y * a.nz + z : (face < 4 ? x * a.nz + z : x * a.ny + y);
    return a.vbc_off[face] + s;
}

// Compute_S_local :278-299
__device__ __forceinline__ void s_local(const Step2Args &A, float psi, float &sv, float &so) {
    if (psi > 0.f)
        sv = psi > 0.1f ? A.wl : A.lg0 + A.l1 * psi + A.l2 * psi * psi;
    else
        sv = psi < -0.1f ? A.wg : A.lg0 + A.g1 * psi + A.g2 * psi * psi;
    so = 8.0f * (2.0f - sv) / (8.0f - sv);
}

#ifdef LBM_STRICT
// ---- literal evaluation order (oracle/ref_two_phase.c) ----------------------------------------
__device__ __forceinline__ void macro2(const float (&f)[19], const LbmParams &P, float &rho, float &ux,
                                       float &uy, float &uz) {
    macro(f, P.force, true, rho, ux, uy, uz);       // same interleaved sums as streaming3 :596-602
}

__device__ __forceinline__ void collide2(float (&f)[19], const Step2Args &A, bool force, float rho, float ux,
                                         float uy, float uz, float psi, float Cx, float Cy, float Cz) {
    const LbmParams &P = A.a.P;
    constexpr int M[19][19] = {
        {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
        {-1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
        {1, -2, -2, -2, -2, -2, -2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
        {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0},
        {0, -2, 2, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0},
        {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1},
        {0, 0, 0, -2, 2, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1},
        {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1},
        {0, 0, 0, 0, 0, -2, 2, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1},
        {0, 2, 2, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
        {0, -2, -2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
        {0, 0, 0, 1, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
        {0, 0, 0, -1, -1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, -1, 1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1}};
    constexpr int EV[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
        {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1},
        {1, 0, -1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};
    const float cc = sqrtf(Cx * Cx + Cy * Cy + Cz * Cz);
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (cc > 0.f) { nx = Cx / cc; ny = Cy / cc; nz = Cz / cc; }
    float m[19], meq[19];
#pragma unroll
    for (int s = 0; s < 19; ++s) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 19; ++l)
            if (M[s][l] != 0) acc = acc + (float)M[s][l] * f[l];
        m[s] = acc;
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) meq[s] = 0.f;
    meq[0] = rho; meq[3] = ux; meq[5] = uy; meq[7] = uz;
    meq[1] = ux * ux + uy * uy + uz * uz;
    meq[9] = 2.0f * ux * ux - uy * uy - uz * uz;
    meq[11] = uy * uy - uz * uz;
    meq[13] = ux * uy; meq[14] = uy * uz; meq[15] = ux * uz;
    meq[1] = meq[1] + A.CapA * cc;                                            // :316-321
    meq[9] = meq[9] + 0.5f * A.CapA * cc * (2.0f * nx * nx - ny * ny - nz * nz);
    meq[11] = meq[11] + 0.5f * A.CapA * cc * (ny * ny - nz * nz);
    meq[13] = meq[13] + 0.5f * A.CapA * cc * (nx * ny);
    meq[14] = meq[14] + 0.5f * A.CapA * cc * (ny * nz);
    meq[15] = meq[15] + 0.5f * A.CapA * cc * (nx * nz);
    float sv, so;
    s_local(A, psi, sv, so);
    const float S[19] = {0.f, sv, sv, 0.f, so, 0.f, so, 0.f, so, sv, sv, sv, sv, sv, sv, sv, so, so, so};
    const float fx = P.force[0], fy = P.force[1], fz = P.force[2];
    (void)force;
#pragma unroll
    for (int s = 0; s < 19; ++s) {                                            // :329-331
        m[s] = m[s] - S[s] * (m[s] - meq[s]);
        float guo = 0.f;
#pragma unroll
        for (int l = 0; l < 19; ++l) {
            if (M[s][l] == 0) continue;
            const float e0 = (float)EV[l][0], e1 = (float)EV[l][1], e2 = (float)EV[l][2];
            const float emu_f = (e0 - ux) * fx + (e1 - uy) * fy + (e2 - uz) * fz;
            const float eu = e0 * ux + e1 * uy + e2 * uz;
            const float ef = e0 * fx + e1 * fy + e2 * fz;
            guo = guo + weight(l) * (emu_f + (eu * ef)) * (float)M[s][l];
        }
        m[s] = m[s] + (1.0f - 0.5f * S[s]) * guo;
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 19; ++l) {
            const float a = c_invM[s * 19 + l];
            if (a != 0.f) acc = acc + a * m[l];
        }
        f[s] = acc;
    }
}
#else
// ---- production arithmetic ---------------------------------------------------------------------
__device__ __forceinline__ void macro2(const float (&f)[19], const LbmParams &P, float &rho, float &ux,
                                       float &uy, float &uz) {
    macro(f, P.force, true, rho, ux, uy, uz);
}

__device__ __forceinline__ void collide2(float (&f)[19], const Step2Args &A, bool force, float rho, float ux,
                                         float uy, float uz, float psi, float Cx, float Cy, float Cz) {
    const LbmParams &P = A.a.P;
    float m[19];
    forward(f, m);
    const float c2 = Cx * Cx + Cy * Cy + Cz * Cz;
    const float cc = sqrtf(c2);
    // 0.5 CapA cc (n_a n_b) = 0.5 CapA C_a C_b / cc
    const float k = cc > 0.f ? 0.5f * A.CapA / cc : 0.f;
    float sv, so;
    s_local(A, psi, sv, so);
    const float uxx = ux * ux, uyy = uy * uy, uzz = uz * uz;
    (void)rho;
    // rows 0,3,5,7 have S = 0 (:293-297)
    m[1] = m[1] - sv * (m[1] - (uxx + uyy + uzz + A.CapA * cc));
    m[2] = m[2] - sv * m[2];
    m[4] = m[4] - so * m[4];
    m[6] = m[6] - so * m[6];
    m[8] = m[8] - so * m[8];
    m[9] = m[9] - sv * (m[9] - (2.0f * uxx - uyy - uzz + k * (2.0f * Cx * Cx - Cy * Cy - Cz * Cz)));
    m[10] = m[10] - sv * m[10];
    m[11] = m[11] - sv * (m[11] - (uyy - uzz + k * (Cy * Cy - Cz * Cz)));
    m[12] = m[12] - sv * m[12];
    m[13] = m[13] - sv * (m[13] - (ux * uy + k * Cx * Cy));
    m[14] = m[14] - sv * (m[14] - (uy * uz + k * Cy * Cz));
    m[15] = m[15] - sv * (m[15] - (ux * uz + k * Cx * Cz));
    m[16] = m[16] - so * m[16];
    m[17] = m[17] - so * m[17];
    m[18] = m[18] - so * m[18];
    if (force) {
        // GuoF :241-247 in closed form (un-scaled): non-zero for s in {0,1,3,5,7,9,11,13,14,15}
        const float fx = P.force[0], fy = P.force[1], fz = P.force[2];
        const float xx = fx * ux, yy = fy * uy, zz = fz * uz;
        const float vf = xx + yy + zz;
        const float hv = 1.0f - 0.5f * sv;
        m[0] += (-2.0f / 3.0f) * vf;
        m[1] += hv * (2.0f / 9.0f) * vf;
        m[3] += fx * (1.0f / 3.0f);
        m[5] += fy * (1.0f / 3.0f);
        m[7] += fz * (1.0f / 3.0f);
        m[9] += hv * (2.0f / 9.0f) * (2.0f * xx - yy - zz);
        m[11] += hv * (2.0f / 9.0f) * (yy - zz);
        m[13] += hv * (1.0f / 9.0f) * (fx * uy + fy * ux);
        m[14] += hv * (1.0f / 9.0f) * (fy * uz + fz * uy);
        m[15] += hv * (1.0f / 9.0f) * (fx * uz + fz * ux);
    }
    inverse(m, f);
}
#endif

// Accumulators of the colour pass: one running sum per colour, every term weighted before it
// is added, directions ascending (the oracle's order; the reference's float atomics have none).
// (Summing unweighted per weight class and weighting once at the end saves 4 of 12 flops per
// direction but no time -- the pass is bound by its 38 gathers per node -- and its different
// rounding flips the |rho_r - rho_b| > 0.9 wetting switch (:271) at a few nodes of config 4.)
struct ColourSum {
    float r = 0.f, b = 0.f;                  // accumulators start at 0 (:596)
    __device__ __forceinline__ float red() const { return r; }
    __device__ __forceinline__ float blue() const { return b; }
};

// Contribution of one pull source to rho_r, rho_b: the recoloured g_r[s], g_b[s] (:345-363) of
// the source node, re-evaluated from its colour record, for the direction sg*e_S (sg = -1:
// the opposite direction LR[S], evaluated on the node's OWN record when the source is solid
// and its push came back, :370-372).  Negation commutes with rounding, min(a,b,c,d) is
// symmetric and cs*(-x) = -(cs*x), so both members of a pair (kk, kk+1) reduce to
//     g_r += cs (e.C)/|C| ,  g_b -= cs (e.C)/|C|     with e the direction being evaluated,
// bit-identically to the reference's pairwise update.
template <int S, int EX, int EY, int EZ>
__device__ __forceinline__ void colour_add(float sg, const float4 ra, const float2 rq, const float4 *__restrict__ pc,
                                           ColourSum &acc) {
    const float eu = sg * edotu<EX, EY, EZ>(ra.z, ra.w, rq.x);
#ifdef LBM_STRICT
    float gr, gb;
    const float uv = ra.z * ra.z + ra.w * ra.w + rq.x * rq.x;
    const float T1 = 1.0f + 3.0f * eu + 4.5f * eu * eu - 1.5f * uv;        // feq :161-170
    gr = weight(S) * ra.x * T1;
    gb = weight(S) * ra.y * T1;
    if (S > 0 && rq.y < 0.f) {                   // flagged: |C| > 0
        const float4 rc = __ldg(pc);
        const float cc = sqrtf(rc.x * rc.x + rc.y * rc.y + rc.z * rc.z);
        const float em = -eu;
        const float T2 = 1.0f + 3.0f * em + 4.5f * em * em - 1.5f * uv;
        const float gro = weight(S) * ra.x * T2, gbo = weight(S) * ra.y * T2;
        float cs = gr < gro ? gr : gro;
        cs = cs < gb ? cs : gb;
        cs = cs < gbo ? cs : gbo;
        cs = cs * ((sg * edotu<EX, EY, EZ>(rc.x, rc.y, rc.z)) / cc);
        gr = gr + cs;
        gb = gb - cs;
    }
    acc.r = acc.r + gr;
    acc.b = acc.b + gb;
#else
    // feq(s) = w rho t with t = q + eu (3 + 4.5 eu); the opposite direction has t - 6 eu
    const float t = fabsf(rq.y) + eu * (3.0f + 4.5f * eu);
    float gr = ra.x * t, gb = ra.y * t;
    if (S > 0 && rq.y < 0.f) {                   // interface node; most nodes skip this
        const float4 rc = __ldg(pc);
        const float to = t - 6.0f * eu;
        const float cs = fminf(fminf(gr, ra.x * to), fminf(gb, ra.y * to)) *
                         (sg * edotu<EX, EY, EZ>(rc.x, rc.y, rc.z) * rc.w);
        gr = gr + cs;
        gb = gb - cs;
    }
    acc.r = acc.r + gr * weight(S);
    acc.b = acc.b + gb * weight(S);
#endif
}

// ---------------------------------------------------------------------------------------------
// colour pass
// ---------------------------------------------------------------------------------------------
#ifndef LBM2P_COLOUR_MINB
#define LBM2P_COLOUR_MINB 8
#endif
#ifndef LBM2P_MAIN_MINB
#define LBM2P_MAIN_MINB 5
#endif
__global__ void __launch_bounds__(256, LBM2P_COLOUR_MINB) k2p_colour(const Step2Args A) {
    const StepArgs &a = A.a;
    // blockIdx.x walks the z-chunks of a row (fastest), (blockIdx.z, blockIdx.y) the row groups:
    // blocks that run together then cover whole z-rows, i.e. contiguous 19*nzp*4-byte chunks
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    if (z >= (uint32_t)a.nz || r >= a.row_count) return;
    const uint32_t row = a.row_first + r;
    const uint32_t idx = row * (uint32_t)a.nz + z;
    const uint8_t cls = a.cls[idx];
    if (cls == NODE_SOLID || cls == NODE_SOLID_WRITE) return;
    const int sx = a.ny * a.nz, sy = a.nz;
    const float4 *__restrict__ pA = A.recA + idx;
    const float2 *__restrict__ pB = A.recB + idx;
    const float4 *__restrict__ pC = A.recC + idx;
    ColourSum acc;
    uint32_t fl = 0;
    if (cls == NODE_BULK) {
        // no solid link, no wrap: sources at uniform offsets
#define X(s, ex, ey, ez, o)                                                                    \
    {                                                                                          \
        const int off = -((ex) * sx + (ey) * sy + (ez));                                       \
        colour_add<s, ex, ey, ez>(1.0f, __ldg(pA + off), __ldg(pB + off), pC + off, acc);   \
    }
        D3Q19_DIRS(X)
#undef X
    } else {
        fl = a.flags[idx];
        // with ghost planes (x-slab) the x neighbours are always at -+sx: no periodic wrap
        const uint32_t flw = a.halo_x ? fl & ~(FL_AT_X0 | FL_AT_X1) : fl;
        // node-linear offsets to x-1 / x+1 ... with the periodic wrap of periodic_index :377-387
        const int oxm = (flw & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
        const int oxp = (flw & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
        const int oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
        const int oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
        const int ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
        const int ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
#define X(s, ex, ey, ez, o)                                                                    \
    {                                                                                          \
        const bool bounce = (fl >> s) & 1u;                                                    \
        const int off = (s == 0 || bounce) ? 0 : OFF(ex, ey, ez);                              \
        colour_add<s, ex, ey, ez>(bounce ? -1.0f : 1.0f, __ldg(pA + off), __ldg(pB + off), pC + off, acc); \
    }
        D3Q19_DIRS(X)
#undef X
#undef OFF
    }
    float rr = acc.red(), rb = acc.blue();
    float psi = rr - rb / (rr + rb);         // :605, precedence as written
    // Boundary_condition_psi :445-486, faces in order, the last matching face wins
    int win = -1;
    if (fl & (FL_AT_X0 | FL_AT_X1 | FL_AT_Y0 | FL_AT_Y1 | FL_AT_Z0 | FL_AT_Z1)) {
        if ((fl & FL_AT_X0) && A.bc_psi_type[0] == 1) win = 0;
        if ((fl & FL_AT_X1) && A.bc_psi_type[1] == 1) win = 1;
        if ((fl & FL_AT_Y0) && A.bc_psi_type[2] == 1) win = 2;
        if ((fl & FL_AT_Y1) && A.bc_psi_type[3] == 1) win = 3;
        if ((fl & FL_AT_Z0) && A.bc_psi_type[4] == 1) win = 4;
        if ((fl & FL_AT_Z1) && A.bc_psi_type[5] == 1) win = 5;
    }
    if (win >= 0) {
        psi = A.bc_psi_val[win];
        rr = (psi + 1.0f) / 2.0f;
        rb = 1.0f - rr;
    }
    A.rho_r[idx] = rr;
    A.rho_b[idx] = rb;
    A.psi[idx] = psi;
}

// ---------------------------------------------------------------------------------------------
// main pass
// ---------------------------------------------------------------------------------------------
template <bool FORCE, int MODE, bool SPEC>
__global__ void __launch_bounds__(256, LBM2P_MAIN_MINB) k2p_main(const Step2Args A) {
    const StepArgs &a = A.a;
    // blockIdx.x walks the z-chunks of a row (fastest), (blockIdx.z, blockIdx.y) the row groups:
    // blocks that run together then cover whole z-rows, i.e. contiguous 19*nzp*4-byte chunks
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    if (z >= (uint32_t)a.nz || r >= a.row_count) return;
    const uint32_t row = a.row_first + r;
    const uint32_t idx = row * (uint32_t)a.nz + z;
    const uint32_t pidx = row * a.prow + z;
    float f[19];
    float rho = 1.0f, ux = 0.f, uy = 0.f, uz = 0.f;
    const uint8_t cls = a.cls[idx];
    if (!SPEC && cls == NODE_SOLID) return;
    uint32_t fl = 0;
    bool compute = true;
    bool pressure = false;
    uint32_t slot = 0;
    if (MODE == MODE_COLLIDE) {
        if (cls == NODE_SOLID || cls == NODE_SOLID_WRITE) return;
        fl = cls == NODE_SPECIAL ? a.flags[idx] : 0u;
        if (a.F != nullptr) {
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = a.F[(size_t)idx * 19 + s];
            rho = a.rho[idx];
            ux = a.v[(size_t)idx * 3 + 0];
            uy = a.v[(size_t)idx * 3 + 1];
            uz = a.v[(size_t)idx * 3 + 2];
        } else {                          // pristine init() :173-186: F = w, rho = 1, v = 0
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = weight(s);
        }
        if (a.has_bc) {
            const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
            if (bc && a.P.bc_type[bc - 1] == 1) {
                slot = vbc_slot2(a, (int)bc - 1, idx);
                a.vbc[3 * (size_t)slot + 0] = ux;
                a.vbc[3 * (size_t)slot + 1] = uy;
                a.vbc[3 * (size_t)slot + 2] = uz;
            }
        }
    } else {
        if (SPEC || cls != NODE_SOLID_WRITE) {
#define X(s, ex, ey, ez, o) f[s] = __ldg(a.ppull[s] + pidx);
            D3Q19_DIRS(X)
#undef X
        } else {
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = 0.f;
        }
        if (SPEC && cls == NODE_SOLID) return;
        compute = cls != NODE_SOLID_WRITE;
        if (MODE == MODE_EXTRACT && !compute) return;
        if (cls == NODE_SPECIAL) {
            fl = a.flags[idx];
            // with ghost planes (x-slab) the x neighbours are always at -+sx: no periodic wrap
            const uint32_t flw = a.halo_x ? fl & ~(FL_AT_X0 | FL_AT_X1) : fl;
            const int sx = a.ny * (int)a.prow, sy = (int)a.prow;
            const int oxm = (flw & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
            const int oxp = (flw & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
            const int oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
            const int oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
            const int ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
            const int ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
#define WRAPS(ex, ey, ez)                                                                      \
    (flw & ((ex > 0 ? FL_AT_X0 : (ex < 0 ? FL_AT_X1 : 0u)) | (ey > 0 ? FL_AT_Y0 : (ey < 0 ? FL_AT_Y1 : 0u)) | \
           (ez > 0 ? FL_AT_Z0 : (ez < 0 ? FL_AT_Z1 : 0u))))
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        if ((fl >> s) & 1u) f[s] = __ldg(a.pown[o] + pidx);                                    \
        else if (WRAPS(ex, ey, ez)) f[s] = __ldg(a.pown[s] + (pidx + OFF(ex, ey, ez)));        \
    }
            D3Q19_DIRS(X)
#undef X
#undef WRAPS
#undef OFF
            if (a.has_bc) {
                // Boundary_condition :491-583, faces in order x0,x1,y0,y1,z0,z1.  A pressure face
                // overwrites all 19 populations, so only the LAST one matters (link word); the
                // velocity form reads the current F, so every velocity face after it is applied
                // in order.
                const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
                int after = 0;
                if (bc) {                             // :493-498  F = feq(rho_bc, v_prev)
                    const int face = (int)bc - 1;
                    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
                    slot = vbc_slot2(a, face, idx);
                    if (!(fl & FL_PIN_SOLID)) {
                        u0 = a.vbc[3 * (size_t)slot + 0];
                        u1 = a.vbc[3 * (size_t)slot + 1];
                        u2 = a.vbc[3 * (size_t)slot + 2];
                    }
                    feq_all(f, a.P.bc_rho[face], u0, u1, u2);
                    pressure = true;
                    after = face + 1;
                }
                for (int face = after; face < 6; ++face) {
                    if (a.P.bc_type[face] != 2 || !(fl & (FL_AT_X0 << face))) continue;
                    // :500-504  F[s] = feq(LR[s],1,u) - F[LR[s]] + feq(s,1,u), IN PLACE for
                    // s = 0..18, u = bc_vel (never written by the reference: zero)
                    const float u0 = a.P.bc_vel[face][0], u1 = a.P.bc_vel[face][1], u2 = a.P.bc_vel[face][2];
#define X(s, ex, ey, ez, o)                                                                    \
    f[s] = feq<o, -(ex), -(ey), -(ez)>(1.0f, u0, u1, u2) - f[o] + feq<s, ex, ey, ez>(1.0f, u0, u1, u2);
                    D3Q19_DIRS(X)
#undef X
                }
            }
        }
        if (compute) macro2(f, a.P, rho, ux, uy, uz);
        if (MODE == MODE_EXTRACT) {
            a.rho[idx] = rho;
            a.v[(size_t)idx * 3 + 0] = ux;
            a.v[(size_t)idx * 3 + 1] = uy;
            a.v[(size_t)idx * 3 + 2] = uz;
            if (a.F != nullptr) {
#pragma unroll
                for (int s = 0; s < 19; ++s) a.F[(size_t)idx * 19 + s] = f[s];
            }
            return;
        }
        if (pressure) {
            a.vbc[3 * (size_t)slot + 0] = ux;
            a.vbc[3 * (size_t)slot + 1] = uy;
            a.vbc[3 * (size_t)slot + 2] = uz;
        }
    }
    if (compute) {
        // Compute_C :259-275: C = sum_s 3 w_s e_s psi(i + e_s); solid nodes of the psi array hold
        // psi_solid; constant-psi faces clamp the stencil (:390-428)
        const int sx = a.ny * a.nz, sy = a.nz;
        int oxm = -sx, oxp = sx, oym = -sy, oyp = sy, ozm = -1, ozp = 1;
        if (fl & (FL_AT_X0 | FL_AT_X1 | FL_AT_Y0 | FL_AT_Y1 | FL_AT_Z0 | FL_AT_Z1)) {
            if (fl & FL_AT_X0) oxm = A.bc_psi_type[0] == 0 ? (a.halo_x ? -sx : (a.nx - 1) * sx) : 0;
            if (fl & FL_AT_X1) oxp = A.bc_psi_type[1] == 0 ? (a.halo_x ? sx : -(a.nx - 1) * sx) : 0;
            if (fl & FL_AT_Y0) oym = A.bc_psi_type[2] == 0 ? (a.ny - 1) * sy : 0;
            if (fl & FL_AT_Y1) oyp = A.bc_psi_type[3] == 0 ? -(a.ny - 1) * sy : 0;
            if (fl & FL_AT_Z0) ozm = A.bc_psi_type[4] == 0 ? (a.nz - 1) : 0;
            if (fl & FL_AT_Z1) ozp = A.bc_psi_type[5] == 0 ? -(a.nz - 1) : 0;
        }
        const float *__restrict__ ps = A.psi + idx;
        float Cx = 0.f, Cy = 0.f, Cz = 0.f;
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        const float val = __ldg(ps + ((ex > 0 ? oxp : (ex < 0 ? oxm : 0)) + (ey > 0 ? oyp : (ey < 0 ? oym : 0)) + \
                                      (ez > 0 ? ozp : (ez < 0 ? ozm : 0))));                   \
        if (ex != 0) Cx = Cx + (3.0f * weight(s) * (float)(ex)) * val;                           \
        if (ey != 0) Cy = Cy + (3.0f * weight(s) * (float)(ey)) * val;                           \
        if (ez != 0) Cz = Cz + (3.0f * weight(s) * (float)(ez)) * val;                           \
    }
        D3Q19_DIRS(X)
#undef X
        const float rr = A.rho_r[idx], rb = A.rho_b[idx];
        if ((fl & FL_NEAR_SOLID) && fabsf(rr - rb) > 0.9f) { Cx = 0.f; Cy = 0.f; Cz = 0.f; }   // :271-273
        const float psi = ps[0];
        collide2(f, A, FORCE, rho, ux, uy, uz, psi, Cx, Cy, Cz);
        // colour record of this collision (see lbm2p_kernels.cuh)
        const float c2 = Cx * Cx + Cy * Cy + Cz * Cz;
        const float ccn = sqrtf(c2);
        const float q = 1.0f - 1.5f * (ux * ux + uy * uy + uz * uz);
        A.recA[idx] = make_float4(rr, rb, ux, uy);
        A.recB[idx] = make_float2(uz, ccn > 0.f ? -q : q);
        if (ccn > 0.f) A.recC[idx] = make_float4(Cx, Cy, Cz, 1.0f / ccn);
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) a.pout[s][pidx] = f[s];
}

// ---------------------------------------------------------------------------------------------
// sparse storage (2phase/lbm_solver_3d_2phase_sparse.py: same arithmetic as the dense script,
// pointer-SNode allocation): compact fluid list + the compressed pull table of the single-phase
// solver.  The node at i + e_s is the pull source of the opposite direction, so one table serves
// the populations (sources i - e_s), the colour records (same sources) and the psi stencil.
// ---------------------------------------------------------------------------------------------
static __constant__ int8_t c_ev[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
    {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1}, {1, 0, -1},
    {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};
static __constant__ int8_t c_lr[19] = {0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17};
// direction index of an offset (cx,cy,cz) in {-1,0,1}^3, -1 if it is not a D3Q19 direction
static __constant__ int8_t c_dir27[27] = {-1, 8, -1, 12, 2, 14, -1, 10, -1, 16, 4, 18, 6, 0, 5, 17, 3, 15,
                                          -1, 9, -1, 13, 1, 11, -1, 7, -1};

// stored index of the pull source of direction s (the node at i - e_s), s a run-time value
__device__ __noinline__ int32_t source_rt(int s, uint32_t i, uint32_t fl, const int32_t (&rb)[8]) {
    switch (s) {
#define X(s_, ex, ey, ez, o) case s_: return s_ == 0 ? (int32_t)i : comp_source<ex, ey, ez>(i, fl, rb);
        D3Q19_DIRS(X)
#undef X
    }
    return (int32_t)i;
}

// psi stencil of Compute_C (:259-275) for a node on a constant-psi face: periodic_index_for_psi
// (:390-428) clamps the out-of-range coordinate, i.e. the neighbour loses that component of e_s
__device__ __noinline__ void gradient_clamped(const Step2Args &A, uint32_t i, uint32_t fl, const int32_t (&rb)[8],
                                              bool exc, uint32_t slot, float &Cx, float &Cy, float &Cz) {
    const StepArgs &a = A.a;
    const float psi_i = A.psi[i];
    Cx = 0.f; Cy = 0.f; Cz = 0.f;
    for (int s = 1; s < 19; ++s) {
        int c[3] = {c_ev[s][0], c_ev[s][1], c_ev[s][2]};
        for (int d = 0; d < 3; ++d) {
            if (c[d] < 0 && (fl & (FL_AT_X0 << (2 * d))) && A.bc_psi_type[2 * d] == 1) c[d] = 0;
            if (c[d] > 0 && (fl & (FL_AT_X1 << (2 * d))) && A.bc_psi_type[2 * d + 1] == 1) c[d] = 0;
        }
        const int sp = c_dir27[(c[0] + 1) * 9 + (c[1] + 1) * 3 + (c[2] + 1)];
        float val = psi_i;
        if (sp > 0) {
            const int o = c_lr[sp];                  // the neighbour at +e_sp is the pull source of o
            if ((fl >> o) & 1u) val = A.psi_solid;
            else val = A.psi[exc ? (uint32_t)__ldg(a.exc[o - 1] + slot) : (uint32_t)source_rt(o, i, fl, rb)];
        }
        const float w3 = 3.0f * weight(s);
        if (c_ev[s][0] != 0) Cx = Cx + (w3 * (float)c_ev[s][0]) * val;
        if (c_ev[s][1] != 0) Cy = Cy + (w3 * (float)c_ev[s][1]) * val;
        if (c_ev[s][2] != 0) Cz = Cz + (w3 * (float)c_ev[s][2]) * val;
    }
}

__global__ void __launch_bounds__(SPARSE_BLOCK, LBM2P_COLOUR_MINB) k2p_colour_sparse(const __grid_constant__ Step2Args A) {
    const StepArgs &a = A.a;
    __shared__ SparseTable s_tab;
    __shared__ uint64_t s_bar;
    const uint32_t blk = blockIdx.x + a.first / SPARSE_BLOCK;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) table_fetch(a, blk, s_tab, &s_bar);
    const uint32_t i = blk * SPARSE_BLOCK + threadIdx.x;
    mbar_wait(&s_bar, 0);
    if (i < a.first || i >= a.first + a.count) return;
    const uint32_t fl = s_tab.fl[threadIdx.x];
    const bool exc = fl & FL_EXCEPTION;
    int32_t rb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rb[k] = s_tab.blk[k] + (int32_t)s_tab.rb[k][threadIdx.x];
    const uint32_t slot = (uint32_t)s_tab.blk[8] + s_tab.rb[0][threadIdx.x];
    ColourSum acc;
#define X(s, ex, ey, ez, o)                                                                    \
    {                                                                                          \
        const bool bounce = s > 0 && ((fl >> s) & 1u);                                         \
        uint32_t j = i;                                                                        \
        if (s > 0 && !bounce) j = exc ? (uint32_t)__ldg(a.exc[s > 0 ? s - 1 : 0] + slot) : (uint32_t)comp_source<ex, ey, ez>(i, fl, rb); \
        colour_add<s, ex, ey, ez>(bounce ? -1.0f : 1.0f, __ldg(A.recA + j), __ldg(A.recB + j), A.recC + j, acc); \
    }
    D3Q19_DIRS(X)
#undef X
    float rr = acc.red(), rbl = acc.blue();
    float psi = rr - rbl / (rr + rbl);       // :605, precedence as written
    // Boundary_condition_psi :445-486, faces in order, the last matching face wins
    int win = -1;
    if (fl & (FL_AT_X0 | FL_AT_X1 | FL_AT_Y0 | FL_AT_Y1 | FL_AT_Z0 | FL_AT_Z1)) {
        if ((fl & FL_AT_X0) && A.bc_psi_type[0] == 1) win = 0;
        if ((fl & FL_AT_X1) && A.bc_psi_type[1] == 1) win = 1;
        if ((fl & FL_AT_Y0) && A.bc_psi_type[2] == 1) win = 2;
        if ((fl & FL_AT_Y1) && A.bc_psi_type[3] == 1) win = 3;
        if ((fl & FL_AT_Z0) && A.bc_psi_type[4] == 1) win = 4;
        if ((fl & FL_AT_Z1) && A.bc_psi_type[5] == 1) win = 5;
    }
    if (win >= 0) {
        psi = A.bc_psi_val[win];
        rr = (psi + 1.0f) / 2.0f;
        rbl = 1.0f - rr;
    }
    A.rho_r[i] = rr;
    A.rho_b[i] = rbl;
    A.psi[i] = psi;
}

template <bool FORCE, int MODE>
__global__ void __launch_bounds__(SPARSE_BLOCK, LBM2P_MAIN_MINB) k2p_main_sparse(const __grid_constant__ Step2Args A) {
    const StepArgs &a = A.a;
    __shared__ SparseTable s_tab;
    __shared__ uint64_t s_bar;
    const uint32_t blk = blockIdx.x + a.first / SPARSE_BLOCK;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) table_fetch(a, blk, s_tab, &s_bar);
    const uint32_t i = blk * SPARSE_BLOCK + threadIdx.x;
    mbar_wait(&s_bar, 0);
    if (i < a.first || i >= a.first + a.count) return;
    const uint32_t fl = s_tab.fl[threadIdx.x];
    const bool exc = fl & FL_EXCEPTION;
    int32_t rb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rb[k] = s_tab.blk[k] + (int32_t)s_tab.rb[k][threadIdx.x];
    const uint32_t slot = (uint32_t)s_tab.blk[8] + s_tab.rb[0][threadIdx.x];
    const bool need_lin = MODE != MODE_STEP || a.has_bc;
    const uint32_t lin = need_lin ? a.lin[i] : 0u;
    float f[19];
    float rho = 1.0f, ux = 0.f, uy = 0.f, uz = 0.f;
    bool pressure = false;
    uint32_t vslot = 0;
    if (MODE == MODE_COLLIDE) {
        if (a.F != nullptr) {
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = a.F[(size_t)lin * 19 + s];
            rho = a.rho[lin];
            ux = a.v[(size_t)lin * 3 + 0];
            uy = a.v[(size_t)lin * 3 + 1];
            uz = a.v[(size_t)lin * 3 + 2];
        } else {                          // pristine init() :173-186: F = w, rho = 1, v = 0
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = weight(s);
        }
        if (a.has_bc) {
            const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
            if (bc && a.P.bc_type[bc - 1] == 1) {
                vslot = vbc_slot2(a, (int)bc - 1, lin);
                a.vbc[3 * (size_t)vslot + 0] = ux;
                a.vbc[3 * (size_t)vslot + 1] = uy;
                a.vbc[3 * (size_t)vslot + 2] = uz;
            }
        }
    } else {
        // pull-stream with half-way bounce-back (:431-443), one selected index per direction
        const int32_t ip = (int32_t)i + (int32_t)a.stride, im = (int32_t)i - (int32_t)a.stride;
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        int32_t j;                                                                             \
        if ((fl >> s) & 1u) j = (o) > (s) ? ip : im;                                           \
        else j = exc ? __ldg(a.exc[s > 0 ? s - 1 : 0] + slot) : comp_source<ex, ey, ez>(i, fl, rb); \
        f[s] = __ldg(a.pown[s] + j);                                                           \
    }
        D3Q19_DIRS(X)
#undef X
        f[0] = __ldg(a.pown[0] + i);
        if (a.has_bc) {
            // Boundary_condition :491-583 (see the dense kernel): last pressure face from the link
            // word, then every later velocity face in order
            const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
            int after = 0;
            if (bc) {
                const int face = (int)bc - 1;
                float u0 = 0.f, u1 = 0.f, u2 = 0.f;
                vslot = vbc_slot2(a, face, lin);
                if (!(fl & FL_PIN_SOLID)) {
                    u0 = a.vbc[3 * (size_t)vslot + 0];
                    u1 = a.vbc[3 * (size_t)vslot + 1];
                    u2 = a.vbc[3 * (size_t)vslot + 2];
                }
                feq_all(f, a.P.bc_rho[face], u0, u1, u2);
                pressure = true;
                after = face + 1;
            }
            for (int face = after; face < 6; ++face) {
                if (a.P.bc_type[face] != 2 || !(fl & (FL_AT_X0 << face))) continue;
                const float u0 = a.P.bc_vel[face][0], u1 = a.P.bc_vel[face][1], u2 = a.P.bc_vel[face][2];
#define X(s, ex, ey, ez, o)                                                                    \
    f[s] = feq<o, -(ex), -(ey), -(ez)>(1.0f, u0, u1, u2) - f[o] + feq<s, ex, ey, ez>(1.0f, u0, u1, u2);
                D3Q19_DIRS(X)
#undef X
            }
        }
        macro2(f, a.P, rho, ux, uy, uz);
        if (MODE == MODE_EXTRACT) {
            a.rho[lin] = rho;
            a.v[(size_t)lin * 3 + 0] = ux;
            a.v[(size_t)lin * 3 + 1] = uy;
            a.v[(size_t)lin * 3 + 2] = uz;
            if (a.F != nullptr) {
#pragma unroll
                for (int s = 0; s < 19; ++s) a.F[(size_t)lin * 19 + s] = f[s];
            }
            return;
        }
        if (pressure) {
            a.vbc[3 * (size_t)vslot + 0] = ux;
            a.vbc[3 * (size_t)vslot + 1] = uy;
            a.vbc[3 * (size_t)vslot + 2] = uz;
        }
    }
    // Compute_C :259-275: C = sum_s 3 w_s e_s psi(i + e_s); a solid neighbour reads psi_solid
    float Cx = 0.f, Cy = 0.f, Cz = 0.f;
    uint32_t clamp = 0;
#pragma unroll
    for (int fc = 0; fc < 6; ++fc)
        if (A.bc_psi_type[fc] == 1) clamp |= FL_AT_X0 << fc;
    if (fl & clamp) {
        gradient_clamped(A, i, fl, rb, exc, slot, Cx, Cy, Cz);
    } else {
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        float val = A.psi_solid;                                                               \
        if (!((fl >> o) & 1u))                                                                 \
            val = __ldg(A.psi + (exc ? (uint32_t)__ldg(a.exc[o > 0 ? o - 1 : 0] + slot)        \
                                     : (uint32_t)comp_source<-(ex), -(ey), -(ez)>(i, fl, rb))); \
        if (ex != 0) Cx = Cx + (3.0f * weight(s) * (float)(ex)) * val;                         \
        if (ey != 0) Cy = Cy + (3.0f * weight(s) * (float)(ey)) * val;                         \
        if (ez != 0) Cz = Cz + (3.0f * weight(s) * (float)(ez)) * val;                         \
    }
        D3Q19_DIRS(X)
#undef X
    }
    const float rr = A.rho_r[i], rbl = A.rho_b[i];
    if ((fl & FL_NEAR_SOLID) && fabsf(rr - rbl) > 0.9f) { Cx = 0.f; Cy = 0.f; Cz = 0.f; }   // :271-273
    const float psi = A.psi[i];
    collide2(f, A, FORCE, rho, ux, uy, uz, psi, Cx, Cy, Cz);
    const float ccn = sqrtf(Cx * Cx + Cy * Cy + Cz * Cz);
    const float q = 1.0f - 1.5f * (ux * ux + uy * uy + uz * uz);
    A.recA[i] = make_float4(rr, rbl, ux, uy);
    A.recB[i] = make_float2(uz, ccn > 0.f ? -q : q);
    if (ccn > 0.f) A.recC[i] = make_float4(Cx, Cy, Cz, 1.0f / ccn);
#pragma unroll
    for (int s = 0; s < 19; ++s) a.pout[s][i] = f[s];
}

cudaError_t launch_main_sparse(int mode, const Step2Args &A, cudaStream_t st) {
    if (A.a.count == 0) return cudaSuccess;
    const unsigned b0 = A.a.first / SPARSE_BLOCK, b1 = (A.a.first + A.a.count + SPARSE_BLOCK - 1) / SPARSE_BLOCK;
    const unsigned grid = b1 - b0;
    switch ((A.a.force ? 4 : 0) | mode) {
        case 0: k2p_main_sparse<false, MODE_STEP><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        case 1: k2p_main_sparse<false, MODE_EXTRACT><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        case 2: k2p_main_sparse<false, MODE_COLLIDE><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        case 4: k2p_main_sparse<true, MODE_STEP><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        case 5: k2p_main_sparse<true, MODE_EXTRACT><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        case 6: k2p_main_sparse<true, MODE_COLLIDE><<<grid, SPARSE_BLOCK, 0, st>>>(A); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_colour_sparse(const Step2Args &A, cudaStream_t st) {
    if (A.a.count == 0) return cudaSuccess;
    const unsigned b0 = A.a.first / SPARSE_BLOCK, b1 = (A.a.first + A.a.count + SPARSE_BLOCK - 1) / SPARSE_BLOCK;
    k2p_colour_sparse<<<b1 - b0, SPARSE_BLOCK, 0, st>>>(A);
    return cudaGetLastError();
}

template <bool FORCE, int MODE>
static void launch_main_t(const Step2Args &A, dim3 grid, dim3 blk, cudaStream_t st) {
    if (A.a.spec)
        k2p_main<FORCE, MODE, true><<<grid, blk, 0, st>>>(A);
    else
        k2p_main<FORCE, MODE, false><<<grid, blk, 0, st>>>(A);
}

// `zmax`: longest z extent of a block.  The main pass streams, so a block is one whole z-row
// (contiguous chunk).  The colour pass gathers 19 neighbour records per node: a block of a
// few y-rows x 64 z shares most of them through L1 (half the L2->L1 traffic of a one-row block).
static void geometry(const StepArgs &a, int block, dim3 &grid, dim3 &blk, int zmax = 256) {
    if (block <= 0 || block > 256 || block % 32) block = 256;
    if (zmax > block) zmax = block;
    // split a z-row into the fewest chunks of at most `zmax` threads, of equal (warp-rounded) size
    const int nchunk = (a.nz + zmax - 1) / zmax;
    int bx = ((a.nz + nchunk - 1) / nchunk + 31) / 32 * 32;
    if (bx > zmax) bx = zmax;
    int by = block / bx;
    if (by < 1) by = 1;
    blk = dim3(bx, by, 1);
    const unsigned rg = (a.row_count + by - 1) / by;
    const unsigned gy = rg < 32768u ? rg : 32768u;
    grid = dim3((a.nz + bx - 1) / bx, gy, (rg + gy - 1) / gy);
}

cudaError_t launch_main(int mode, const Step2Args &A, int block, cudaStream_t st) {
    if (A.a.row_count == 0) return cudaSuccess;
    dim3 grid, blk;
    geometry(A.a, block, grid, blk);
    switch ((A.a.force ? 4 : 0) | mode) {
        case 0: launch_main_t<false, MODE_STEP>(A, grid, blk, st); break;
        case 1: launch_main_t<false, MODE_EXTRACT>(A, grid, blk, st); break;
        case 2: launch_main_t<false, MODE_COLLIDE>(A, grid, blk, st); break;
        case 4: launch_main_t<true, MODE_STEP>(A, grid, blk, st); break;
        case 5: launch_main_t<true, MODE_EXTRACT>(A, grid, blk, st); break;
        case 6: launch_main_t<true, MODE_COLLIDE>(A, grid, blk, st); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_colour(const Step2Args &A, int block, cudaStream_t st) {
    if (A.a.row_count == 0) return cudaSuccess;
    dim3 grid, blk;
    static int zmax = 0;
    if (zmax == 0) {
        const char *e = getenv("LBM3D_COLOUR_BX");       // tuning knob
        zmax = e ? atoi(e) : 32;
        if (zmax < 32 || zmax > 256 || zmax % 32) zmax = 32;
    }
    geometry(A.a, block, grid, blk, zmax);
    k2p_colour<<<grid, blk, 0, st>>>(A);
    return cudaGetLastError();
}

cudaError_t set_inverse_matrix(const float *invM361) {
#ifdef LBM_STRICT
    return cudaMemcpyToSymbol(c_invM, invM361, 361 * sizeof(float));
#else
    (void)invM361;
    return cudaSuccess;
#endif
}

}  // namespace LBM2P_NS
