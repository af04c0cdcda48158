// Two-phase colour-gradient D3Q19 MRT kernels for sm_100a (dense storage).
//
// Reference: 2phase/lbm_solver_3d_2phase.py, main loop :626-632
//     colission :302-372 ; streaming1 :431 ; Boundary_condition :491 ; streaming3 :587 ;
//     Boundary_condition_psi :445
// Two launches per step instead of five passes:
//
//   k2p_colour   rho_r, rho_b of step n from the colour records of the 18 pull sources
//                (the recoloured g_r, g_b of :345-363 are RE-EVALUATED from the source node's
//                record (rho_r, rho_b, v, C) instead of being stored: 6 words per node instead
//                of 38, and a deterministic ascending-s sum instead of the reference's
//                unordered float atomics :365-372); then psi (:605) and the psi BC (:445-486).
//                A thread fetches the records of its 9 neighbour z-rows at its OWN z only; the
//                contributions that travel along z come from the adjacent lanes (warp shuffles,
//                shared memory across warps): 9 aligned record loads per node instead of 19
//                misaligned gathers.
//   k2p_main     pull-stream f* + flow BCs + macro of step n, then the collision of step n+1:
//                C = grad(psi) over 18 neighbours (:259-275), surface-tension perturbation of
//                meq (:316-321), psi-dependent relaxation (:278-299), Guo force (:241-247,
//                closed form), inv_M; writes f* and the node's colour record.
//
// The pull needs psi of step n at all neighbours before C can be formed, hence two kernels.
// Algorithmic bytes per node-step: 152 (populations) + 16 (rho_r, rho_b r/w) +
// 8 (psi r/w) = 176; this implementation moves 176 + 32 (record write + read) = 208.
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
#ifdef LBM_STRICT
    if (psi > 0.f)
        sv = psi > 0.1f ? A.wl : A.lg0 + A.l1 * psi + A.l2 * psi * psi;
    else
        sv = psi < -0.1f ? A.wg : A.lg0 + A.g1 * psi + A.g2 * psi * psi;
    so = 8.0f * (2.0f - sv) / (8.0f - sv);
#else
    // the same piecewise rate without branches; the quotient through the reciprocal unit
    const bool pos = psi > 0.f;
    const float c1 = pos ? A.l1 : A.g1, c2 = pos ? A.l2 : A.g2;
    const float quad = A.lg0 + (c1 + c2 * psi) * psi;
    sv = fabsf(psi) > 0.1f ? (pos ? A.wl : A.wg) : quad;
#ifdef LBM2P_EXACT_DIV
    so = 8.0f * (2.0f - sv) / (8.0f - sv);
#else
    so = __fdividef(8.0f * (2.0f - sv), 8.0f - sv);
#endif
#endif
}

#ifdef LBM_STRICT
// ---- literal evaluation order (oracle/ref_two_phase.c) ----------------------------------------
__device__ __forceinline__ void macro2(const float (&f)[19], const LbmParams &P, float &rho, float &ux,
                                       float &uy, float &uz) {
    macro(f, P.force, true, rho, ux, uy, uz);       // same interleaved sums as streaming3 :596-602
}

__device__ __forceinline__ void collide2(float (&f)[19], const Step2Args &A, bool force, float rho, float ux,
                                         float uy, float uz, float psi, float Cx, float Cy, float Cz, float &inv) {
    const LbmParams &P = A.a.P;
    inv = 0.f;
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
                                         float uy, float uz, float psi, float Cx, float Cy, float Cz, float &inv) {
    const LbmParams &P = A.a.P;
    float m[19];
    forward(f, m);
    const float c2 = Cx * Cx + Cy * Cy + Cz * Cz;
#ifdef LBM2P_EXACT_DIV
    inv = c2 > 0.f ? 1.0f / sqrtf(c2) : 0.f;
#else
    inv = c2 > 0.f ? rsqrtf(c2) : 0.f;       // 1/|C|, shared with the colour record (unit normal)
#endif
    const float cc = c2 * inv;
    // 0.5 CapA cc (n_a n_b) = 0.5 CapA C_a C_b / cc
    const float k = 0.5f * A.CapA * inv;
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

// colour record of a collision (lbm2p_kernels.cuh): velocity, q = 1 - 1.5 v.v with the interface
// flag in its sign, and for flagged nodes the interface vector the colour pass needs
__device__ __forceinline__ void write_record(const Step2Args &A, uint32_t node, float ux, float uy, float uz,
                                             float Cx, float Cy, float Cz, float inv) {
    const float q = 1.0f - 1.5f * (ux * ux + uy * uy + uz * uz);
#ifdef LBM_STRICT
    (void)inv;
    const float ccn = sqrtf(Cx * Cx + Cy * Cy + Cz * Cz);
    A.uq[node] = make_float4(ux, uy, uz, ccn > 0.f ? -q : q);
    if (ccn > 0.f) A.recC[node] = make_float4(Cx, Cy, Cz, 1.0f / ccn);
#else
    A.uq[node] = make_float4(ux, uy, uz, inv > 0.f ? -q : q);
    if (inv > 0.f) A.recC[node] = make_float4(Cx * inv, Cy * inv, Cz * inv, 0.f);
#endif
}

// Accumulators of the colour pass: one running sum per colour, every term weighted before it
// is added, directions ascending (the oracle's order; the reference's float atomics have none).
// (Summing unweighted per weight class and weighting once at the end saves 4 of 12 flops per
// direction but its different rounding flips the |rho_r - rho_b| > 0.9 wetting switch (:271) at a
// few nodes of config 4.)
struct ColourSum {
    float r = 0.f, b = 0.f;                  // accumulators start at 0 (:596)
};

// Contribution of one pull source to rho_r, rho_b: the recoloured g_r[s], g_b[s] (:345-363) of
// the source node, re-evaluated from its colour record (ab = rho_r, rho_b; uq = v, +-q; *pc = C),
// for the direction sg*e_S (sg = -1: the opposite direction LR[S], evaluated on the node's OWN
// record when the source is solid and its push came back, :370-372).  Negation commutes with
// rounding, min(a,b,c,d) is symmetric and cs*(-x) = -(cs*x), so both members of a pair (kk, kk+1)
// reduce to
//     g_r += cs (e.C)/|C| ,  g_b -= cs (e.C)/|C|     with e the direction being evaluated,
// bit-identically to the reference's pairwise update.  Production arithmetic returns the terms
// un-weighted (colour_acc weights them) and spells every rounding out with intrinsics: the term is
// evaluated on the CONSUMER's lane for nodes next to solids / faces and on the SOURCE's lane
// otherwise, and which of the two a node takes depends on the decomposition into slabs.
// `rc` is the interface part of the record, present (and loaded by the caller) only for flagged
// records, uq.w < 0, zero otherwise.  Verification arithmetic: (C, 1/|C|) as the collision used
// them.  Production arithmetic: the unit normal C/|C|, and the recolouring evaluated for every
// record without a branch (a zero normal makes it vanish) -- a few steps into a run psi is exactly
// uniform only in the bulk of the blue phase, so nearly every warp would take the branch anyway:
// with t = q + eu (3 + 4.5 eu) the opposite direction has to = t - 6 eu, and the
// four-way min of :351-356 is min(rho_r x, rho_b y) with x, y the smaller of (t, to) for a
// non-negative density and the larger for a negative one (rounding is monotonic, so this is the
// min of the four rounded products, bit for bit, while t, to > 0).
// the red / blue pairs of a term (two products, two weighted accumulations) as packed fp32 instructions:
// 72 SASS instructions less in the row-sharing colour kernel, 2 % of a step on the 256^3 droplet
#ifndef LBM2P_PACKED_F32
#define LBM2P_PACKED_F32 1
#endif
#if LBM2P_PACKED_F32 && !defined(LBM_STRICT)
// sm_100 packed fp32: two IEEE round-to-nearest operations in one issue slot (same bits as two scalar ones)
__device__ __forceinline__ void fma2(float &r0, float &r1, float a0, float a1, float b0, float b1) {
    unsigned long long A, B, C;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(r0), "f"(r1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(C));
}
__device__ __forceinline__ void mul2(float &r0, float &r1, float a0, float a1, float b0, float b1) {
    unsigned long long A, B, C;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(C));
}
#endif
template <int S, int EX, int EY, int EZ>
__device__ __forceinline__ void colour_term(float sg, const float2 ab, const float4 uq, const float4 rc,
                                            float &gr, float &gb) {
    const float eu = sg * edotu<EX, EY, EZ>(uq.x, uq.y, uq.z);
#ifdef LBM_STRICT
    const float uv = uq.x * uq.x + uq.y * uq.y + uq.z * uq.z;
    const float T1 = 1.0f + 3.0f * eu + 4.5f * eu * eu - 1.5f * uv;        // feq :161-170
    gr = weight(S) * ab.x * T1;
    gb = weight(S) * ab.y * T1;
    if (S > 0 && uq.w < 0.f) {                   // flagged: |C| > 0
        const float cc = sqrtf(rc.x * rc.x + rc.y * rc.y + rc.z * rc.z);
        const float em = -eu;
        const float T2 = 1.0f + 3.0f * em + 4.5f * em * em - 1.5f * uv;
        const float gro = weight(S) * ab.x * T2, gbo = weight(S) * ab.y * T2;
        float cs = gr < gro ? gr : gro;
        cs = cs < gb ? cs : gb;
        cs = cs < gbo ? cs : gbo;
        cs = cs * ((sg * edotu<EX, EY, EZ>(rc.x, rc.y, rc.z)) / cc);
        gr = gr + cs;
        gb = gb - cs;
    }
#else
    const float t = __fmaf_rn(eu, __fmaf_rn(4.5f, eu, 3.0f), fabsf(uq.w));
    if (S > 0) {
        const float to = __fmaf_rn(-6.0f, eu, t);
        const float lo = fminf(t, to), hi = fmaxf(t, to);
#if LBM2P_PACKED_F32
        float pa, pb;
        mul2(pa, pb, ab.x, ab.y, ab.x >= 0.f ? lo : hi, ab.y >= 0.f ? lo : hi);
#else
        const float pa = __fmul_rn(ab.x, ab.x >= 0.f ? lo : hi), pb = __fmul_rn(ab.y, ab.y >= 0.f ? lo : hi);
#endif
        const float cs = __fmul_rn(fminf(pa, pb), sg * edotu<EX, EY, EZ>(rc.x, rc.y, rc.z));
        gr = __fmaf_rn(ab.x, t, cs);
        gb = __fmaf_rn(ab.y, t, -cs);
    } else {
        gr = __fmul_rn(ab.x, t);
        gb = __fmul_rn(ab.y, t);
    }
#endif
}
// interface part of a record (see colour_term): stored, and fetched, only where the collision saw
// C != 0 -- the sign of the record's q says so; a predicated load, no branch
__device__ __forceinline__ float4 load_interface(const float4 *__restrict__ pc, float qflag) {
    return qflag < 0.f ? __ldg(pc) : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int S>
__device__ __forceinline__ void colour_acc(ColourSum &acc, float gr, float gb) {
#ifdef LBM_STRICT
    acc.r = acc.r + gr;
    acc.b = acc.b + gb;
#else
#if LBM2P_PACKED_F32
    fma2(acc.r, acc.b, gr, gb, weight(S), weight(S));
#else
    acc.r = __fmaf_rn(gr, weight(S), acc.r);
    acc.b = __fmaf_rn(gb, weight(S), acc.b);
#endif
#endif
}

// ---------------------------------------------------------------------------------------------
// colour pass
// ---------------------------------------------------------------------------------------------
#ifndef LBM2P_COLOUR_MINB
#define LBM2P_COLOUR_MINB 8
#endif
#ifndef LBM2P_COLOUR_MINB_ROWS
#define LBM2P_COLOUR_MINB_ROWS 5      // 48 registers, no spills (6: 40 registers, 20 B of spills, no faster)
#endif
#ifndef LBM2P_COLOUR_ROWS
#define LBM2P_COLOUR_ROWS 8           // y-rows (warps) of a block of the row-sharing colour kernel (256^3 droplet: 4 -> 1.048, 8 -> 1.042, 16 -> 1.128 ms per step)
#endif
#ifndef LBM2P_MAIN_MINB
#define LBM2P_MAIN_MINB 5             // 48 registers (4: 62 registers, 1.078 instead of 1.042 ms per step)
#endif
#ifndef LBM2P_MAIN_SPARSE_MINB
#define LBM2P_MAIN_SPARSE_MINB 6      // 40 registers; 384^3 pack, ms per step by blocks per SM: 3 -> 1.628, 4 -> 1.384, 5 -> 1.259, 6 -> 1.207, 8 -> 1.426
#endif
// psi (:605) and Boundary_condition_psi (:445-486) from the colour sums; stores the node's state
__device__ __forceinline__ void colour_finish(const Step2Args &A, uint32_t node, uint32_t fl, float rr, float rb) {
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
    A.rrb_out[node] = make_float2(rr, rb);
    A.psi[node] = psi;
}

// GATHER = false (lattices that are mostly bulk fluid): a WARP owns 30 consecutive nodes of a z-row
// and sits on 32 -- one more on either side, wrapped periodically like periodic_index (:377-387).
// Every lane loads the records of its 9 neighbour z-rows at its own z, evaluates on them the
// terms that go to z (own), z+1 (up) and z-1 (dn), and hands the up / dn terms to the adjacent
// lanes with shuffles; the two outer lanes only give.  A node without solid links is complete with
// that.  9 aligned-row record loads per node instead of 19 gathers, and the 8 warps of a block
// (8 consecutive y-rows) find most of their rows in L1.  Every other fluid node (solid link) -- and
// every fluid node when GATHER = true (porous media, where almost every node has a solid link) --
// evaluates its 19 terms itself from gathered records.
#define COLOUR_TILE 30
// Order in which a node adds its 19 terms.  Verification arithmetic: ascending direction, the order
// of the oracle.  Production arithmetic: row by row as the warp tiles produce them -- (own, up, dn)
// of the five axis rows, then the four diagonal rows -- so that a handed-over term is added as soon
// as it arrives instead of waiting in a register for its turn (48 registers instead of 64: 5 blocks
// of 8 warps per SM instead of 4; worth 2 % of a step on the 256^3 droplet, 1.065 -> 1.045 ms -- the
// pass is held back by the L1 data path as much as by latency); the gather paths below and the sparse
// kernel use the same order, which keeps the three bit-identical to each other.
#ifndef LBM2P_COLOUR_TILE_ORDER
#define LBM2P_COLOUR_TILE_ORDER 1
#endif
#if defined(LBM_STRICT) || !LBM2P_COLOUR_TILE_ORDER
#define COLOUR_ROW_ORDER 0
#define D3Q19_DIRS_COLOUR(X) D3Q19_DIRS(X)
#else
#define COLOUR_ROW_ORDER 1
#define D3Q19_DIRS_COLOUR(X)                                                               \
    X(0, 0, 0, 0, 0) X(5, 0, 0, 1, 6) X(6, 0, 0, -1, 5)                                    \
    X(1, 1, 0, 0, 2) X(11, 1, 0, 1, 12) X(13, 1, 0, -1, 14)                                \
    X(2, -1, 0, 0, 1) X(14, -1, 0, 1, 13) X(12, -1, 0, -1, 11)                             \
    X(3, 0, 1, 0, 4) X(15, 0, 1, 1, 16) X(17, 0, 1, -1, 18)                                \
    X(4, 0, -1, 0, 3) X(18, 0, -1, 1, 17) X(16, 0, -1, -1, 15)                             \
    X(7, 1, 1, 0, 8) X(8, -1, -1, 0, 7) X(9, 1, -1, 0, 10) X(10, -1, 1, 0, 9)
#endif
template <bool GATHER>
__global__ void __launch_bounds__(GATHER ? 256 : 32 * LBM2P_COLOUR_ROWS, GATHER ? LBM2P_COLOUR_MINB : LBM2P_COLOUR_MINB_ROWS) k2p_colour(const Step2Args A) {
    const StepArgs &a = A.a;
    // blockIdx.x walks the z-chunks of a row (fastest), (blockIdx.z, blockIdx.y) the row groups
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    const uint32_t nz = (uint32_t)a.nz;
    uint32_t z;
    bool inside;
    if (GATHER) {
        z = blockIdx.x * blockDim.x + threadIdx.x;
        inside = z < nz && r < a.row_count;
        if (!inside) return;
    } else {
        // blockDim.x = 32: lane l of tile t sits on z = 30 t - 1 + l (mod nz) and owns it for l = 1..30
        const int zu = (int)(blockIdx.x * COLOUR_TILE + threadIdx.x) - 1;
        inside = threadIdx.x >= 1u && threadIdx.x <= COLOUR_TILE && zu < (int)nz && r < a.row_count;
        const int zw = zu % (int)nz;
        z = (uint32_t)(zw < 0 ? zw + (int)nz : zw);
    }
    // rows outside the launch stay (the hand-over is warp-wide) on a clamped address
    const uint32_t rc = r < a.row_count ? r : a.row_count - 1u;
    const uint32_t row = a.row_first + rc + (rc >= a.row_split ? a.row_skip : 0u);
    const uint32_t idx = row * nz + z;
    const uint8_t cls = a.cls[idx];
    const bool fluid = inside && (cls == NODE_BULK || cls == NODE_SPECIAL);
    if (GATHER && !fluid) return;
    const uint32_t fl = (fluid && cls == NODE_SPECIAL) ? a.flags[idx] : 0u;
    ColourSum acc;
    bool done = false;
    if (!GATHER) {
        const uint32_t ny = (uint32_t)a.ny, nx = (uint32_t)a.nx;
        const uint32_t x = row / ny, y = row - x * ny;
        // neighbour rows with the periodic wrap (in an x-slab the ghost planes make x +- 1 exist)
        const uint32_t xm = x > 0 ? x - 1 : nx - 1, xp = x + 1 < nx ? x + 1 : 0;
        const uint32_t ym = y > 0 ? y - 1 : ny - 1, yp = y + 1 < ny ? y + 1 : 0;
#if COLOUR_ROW_ORDER
        // axis rows: own term (e_z = 0), up term (e_z = +1, for the node above), dn term (e_z = -1)
#define AXIS_ROW(k, X_, Y_, SO, SU, SD, ex, ey)                                                \
    {                                                                                          \
        const uint32_t nb = ((X_) * ny + (Y_)) * nz + z;                                       \
        const float4 q4 = __ldg(A.uq + nb);                                                    \
        const float2 ab = __ldg(A.rrb + nb);                                                   \
        const float4 n4 = load_interface(A.recC + nb, q4.w);                                   \
        float gr, gb, ur, ub, dr, db;                                                          \
        colour_term<SO, ex, ey, 0>(1.0f, ab, q4, n4, gr, gb);                                  \
        colour_term<SU, ex, ey, 1>(1.0f, ab, q4, n4, ur, ub);                                  \
        colour_term<SD, ex, ey, -1>(1.0f, ab, q4, n4, dr, db);                                 \
        colour_acc<SO>(acc, gr, gb);                                                           \
        colour_acc<SU>(acc, __shfl_up_sync(0xffffffffu, ur, 1), __shfl_up_sync(0xffffffffu, ub, 1));     \
        colour_acc<SD>(acc, __shfl_down_sync(0xffffffffu, dr, 1), __shfl_down_sync(0xffffffffu, db, 1)); \
    }
#else
        float up[5][2], dn[5][2];
        // axis rows: own term (e_z = 0), up term (e_z = +1, for the node above), dn term (e_z = -1)
#define AXIS_ROW(k, X_, Y_, SO, SU, SD, ex, ey)                                                \
    {                                                                                          \
        const uint32_t nb = ((X_) * ny + (Y_)) * nz + z;                                       \
        const float4 q4 = __ldg(A.uq + nb);                                                    \
        const float2 ab = __ldg(A.rrb + nb);                                                   \
        const float4 n4 = load_interface(A.recC + nb, q4.w);                                   \
        float gr, gb, ur, ub, dr, db;                                                          \
        colour_term<SO, ex, ey, 0>(1.0f, ab, q4, n4, gr, gb);                                  \
        colour_term<SU, ex, ey, 1>(1.0f, ab, q4, n4, ur, ub);                                  \
        colour_term<SD, ex, ey, -1>(1.0f, ab, q4, n4, dr, db);                                 \
        colour_acc<SO>(acc, gr, gb);                                                           \
        up[k][0] = __shfl_up_sync(0xffffffffu, ur, 1);                                         \
        up[k][1] = __shfl_up_sync(0xffffffffu, ub, 1);                                         \
        dn[k][0] = __shfl_down_sync(0xffffffffu, dr, 1);                                       \
        dn[k][1] = __shfl_down_sync(0xffffffffu, db, 1);                                       \
    }
#endif
        AXIS_ROW(0, x, y, 0, 5, 6, 0, 0)           // s = 0 ; 5 = (0,0,1) ; 6 = (0,0,-1)
        AXIS_ROW(1, xm, y, 1, 11, 13, 1, 0)        // s = 1 = (1,0,0) ; 11 = (1,0,1) ; 13 = (1,0,-1)
        AXIS_ROW(2, xp, y, 2, 14, 12, -1, 0)       // s = 2 ; 14 = (-1,0,1) ; 12 = (-1,0,-1)
        AXIS_ROW(3, x, ym, 3, 15, 17, 0, 1)        // s = 3 ; 15 = (0,1,1) ; 17 = (0,1,-1)
        AXIS_ROW(4, x, yp, 4, 18, 16, 0, -1)       // s = 4 ; 18 = (0,-1,1) ; 16 = (0,-1,-1)
#undef AXIS_ROW
#if !COLOUR_ROW_ORDER
        colour_acc<5>(acc, up[0][0], up[0][1]);
        colour_acc<6>(acc, dn[0][0], dn[0][1]);
#endif
#define DIAG_ROW(S, X_, Y_, ex, ey)                                                            \
    {                                                                                          \
        const uint32_t nb = ((X_) * ny + (Y_)) * nz + z;                                       \
        const float4 q4 = __ldg(A.uq + nb);                                                    \
        float gr, gb;                                                                          \
        colour_term<S, ex, ey, 0>(1.0f, __ldg(A.rrb + nb), q4, load_interface(A.recC + nb, q4.w), gr, gb); \
        colour_acc<S>(acc, gr, gb);                                                            \
    }
        DIAG_ROW(7, xm, ym, 1, 1)
        DIAG_ROW(8, xp, yp, -1, -1)
        DIAG_ROW(9, xm, yp, 1, -1)
        DIAG_ROW(10, xp, ym, -1, 1)
#undef DIAG_ROW
#if !COLOUR_ROW_ORDER
        colour_acc<11>(acc, up[1][0], up[1][1]);
        colour_acc<12>(acc, dn[2][0], dn[2][1]);
        colour_acc<13>(acc, dn[1][0], dn[1][1]);
        colour_acc<14>(acc, up[2][0], up[2][1]);
        colour_acc<15>(acc, up[3][0], up[3][1]);
        colour_acc<16>(acc, dn[4][0], dn[4][1]);
        colour_acc<17>(acc, dn[3][0], dn[3][1]);
        colour_acc<18>(acc, up[4][0], up[4][1]);
#endif
        if (!fluid) return;
        done = (fl & FL_LINK_MASK) == 0u;          // no solid link: nothing bounced back
    }
    if (!done) {
        acc = ColourSum();
        const int sx = a.ny * a.nz, sy = a.nz;
        // with ghost planes (x-slab) the x neighbours are always at -+sx: no periodic wrap
        const uint32_t flw = a.halo_x ? fl & ~(FL_AT_X0 | FL_AT_X1) : fl;
        // node-linear offsets to x-1 / x+1 ... with the periodic wrap of periodic_index :377-387
        const int oxm = (flw & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
        const int oxp = (flw & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
        const int oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
        const int oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
        const int ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
        const int ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
        const float4 *__restrict__ pU = A.uq + idx;
        const float2 *__restrict__ pR = A.rrb + idx;
        const float4 *__restrict__ pC = A.recC + idx;
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
#define X(s, ex, ey, ez, o)                                                                    \
    {                                                                                          \
        const bool bounce = (fl >> s) & 1u;                                                    \
        const int off = (s == 0 || bounce) ? 0 : OFF(ex, ey, ez);                              \
        float gr, gb;                                                                          \
        const float4 q4 = __ldg(pU + off);                                                     \
        colour_term<s, ex, ey, ez>(bounce ? -1.0f : 1.0f, __ldg(pR + off), q4, load_interface(pC + off, q4.w), gr, gb); \
        colour_acc<s>(acc, gr, gb);                                                            \
    }
        D3Q19_DIRS_COLOUR(X)
#undef X
#undef OFF
    }
    colour_finish(A, idx, fl, acc.r, acc.b);
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
        float Cx = 0.f, Cy = 0.f, Cz = 0.f;
        const float *__restrict__ ps = A.psi + idx;
#define GRAD(val, s, ex, ey, ez)                                                               \
        if (ex != 0) Cx = Cx + (3.0f * weight(s) * (float)(ex)) * val;                           \
        if (ey != 0) Cy = Cy + (3.0f * weight(s) * (float)(ey)) * val;                           \
        if (ez != 0) Cz = Cz + (3.0f * weight(s) * (float)(ez)) * val;
        if (!(fl & (FL_AT_X0 | FL_AT_X1 | FL_AT_Y0 | FL_AT_Y1 | FL_AT_Z0 | FL_AT_Z1))) {
            // away from the lattice faces the stencil sits at fixed offsets
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        const float val = __ldg(A.psi_nb[s] + idx);                                            \
        GRAD(val, s, ex, ey, ez)                                                               \
    }
            D3Q19_DIRS(X)
#undef X
        } else {
            const int sx = a.ny * a.nz, sy = a.nz;
            int oxm = -sx, oxp = sx, oym = -sy, oyp = sy, ozm = -1, ozp = 1;
            if (fl & FL_AT_X0) oxm = A.bc_psi_type[0] == 0 ? (a.halo_x ? -sx : (a.nx - 1) * sx) : 0;
            if (fl & FL_AT_X1) oxp = A.bc_psi_type[1] == 0 ? (a.halo_x ? sx : -(a.nx - 1) * sx) : 0;
            if (fl & FL_AT_Y0) oym = A.bc_psi_type[2] == 0 ? (a.ny - 1) * sy : 0;
            if (fl & FL_AT_Y1) oyp = A.bc_psi_type[3] == 0 ? -(a.ny - 1) * sy : 0;
            if (fl & FL_AT_Z0) ozm = A.bc_psi_type[4] == 0 ? (a.nz - 1) : 0;
            if (fl & FL_AT_Z1) ozp = A.bc_psi_type[5] == 0 ? -(a.nz - 1) : 0;
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        const float val = __ldg(ps + ((ex > 0 ? oxp : (ex < 0 ? oxm : 0)) + (ey > 0 ? oyp : (ey < 0 ? oym : 0)) + \
                                      (ez > 0 ? ozp : (ez < 0 ? ozm : 0))));                   \
        GRAD(val, s, ex, ey, ez)                                                               \
    }
            D3Q19_DIRS(X)
#undef X
        }
#undef GRAD
        if (fl & FL_NEAR_SOLID) {                                   // :271-273 (wetting switch)
            const float2 ab = A.rrb[idx];
            if (fabsf(ab.x - ab.y) > 0.9f) { Cx = 0.f; Cy = 0.f; Cz = 0.f; }
        }
        const float psi = ps[0];
        float inv;
        collide2(f, A, FORCE, rho, ux, uy, uz, psi, Cx, Cy, Cz, inv);
        // colour record of this collision (lbm2p_kernels.cuh); rho_r, rho_b stay where the colour
        // pass left them
        write_record(A, idx, ux, uy, uz, Cx, Cy, Cz, inv);
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
    table_wait(a, blk, s_tab, &s_bar);
    if (i < a.first || i >= a.first + a.count) return;
    const uint32_t fl = s_tab.fl[threadIdx.x];
    const bool exc = fl & FL_EXCEPTION;
    int32_t rb[8];
    table_ranks(s_tab, i, rb);
    const uint32_t slot = table_exc_slot(s_tab);
    ColourSum acc;
#define X(s, ex, ey, ez, o)                                                                    \
    {                                                                                          \
        const bool bounce = s > 0 && ((fl >> s) & 1u);                                         \
        uint32_t j = i;                                                                        \
        if (s > 0 && !bounce) j = exc ? (uint32_t)__ldg(a.exc[s > 0 ? s - 1 : 0] + slot) : (uint32_t)comp_source<ex, ey, ez>(i, fl, rb); \
        float gr, gb;                                                                          \
        const float4 q4 = __ldg(A.uq + j);                                                     \
        colour_term<s, ex, ey, ez>(bounce ? -1.0f : 1.0f, __ldg(A.rrb + j), q4, load_interface(A.recC + j, q4.w), gr, gb); \
        colour_acc<s>(acc, gr, gb);                                                            \
    }
    D3Q19_DIRS_COLOUR(X)
#undef X
    colour_finish(A, i, fl, acc.r, acc.b);
}

template <bool FORCE, int MODE>
__global__ void __launch_bounds__(SPARSE_BLOCK, LBM2P_MAIN_SPARSE_MINB) k2p_main_sparse(const __grid_constant__ Step2Args A) {
    const StepArgs &a = A.a;
    __shared__ SparseTable s_tab;
    __shared__ uint64_t s_bar;
    const uint32_t blk = blockIdx.x + a.first / SPARSE_BLOCK;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) table_fetch(a, blk, s_tab, &s_bar);
    const uint32_t i = blk * SPARSE_BLOCK + threadIdx.x;
    table_wait(a, blk, s_tab, &s_bar);
    if (i < a.first || i >= a.first + a.count) return;
    const uint32_t fl = s_tab.fl[threadIdx.x];
    const bool exc = fl & FL_EXCEPTION;
    int32_t rb[8];
    table_ranks(s_tab, i, rb);
    const uint32_t slot = table_exc_slot(s_tab);
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
    if (fl & FL_NEAR_SOLID) {                                       // :271-273 (wetting switch)
        const float2 ab = A.rrb[i];
        if (fabsf(ab.x - ab.y) > 0.9f) { Cx = 0.f; Cy = 0.f; Cz = 0.f; }
    }
    const float psi = A.psi[i];
    float inv;
    collide2(f, A, FORCE, rho, ux, uy, uz, psi, Cx, Cy, Cz, inv);
    write_record(A, i, ux, uy, uz, Cx, Cy, Cz, inv);
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
    if (!A.gather) {
        // mostly bulk fluid: a warp per 30 nodes of a z-row (+ one either side), 8 y-rows per block
        const unsigned by = LBM2P_COLOUR_ROWS, rg = (A.a.row_count + by - 1) / by, gy = rg < 32768u ? rg : 32768u;
        blk = dim3(32, by, 1);
        grid = dim3((A.a.nz + COLOUR_TILE - 1) / COLOUR_TILE, gy, (rg + gy - 1) / gy);
        k2p_colour<false><<<grid, blk, 0, st>>>(A);
    } else {
        // porous medium: almost every node gathers; a block of a few y-rows x 32 z shares most of
        // the records through L1 (half the L2->L1 traffic of a one-row block)
        geometry(A.a, block, grid, blk, zmax);
        k2p_colour<true><<<grid, blk, 0, st>>>(A);
    }
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
