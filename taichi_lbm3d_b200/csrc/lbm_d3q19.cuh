// D3Q19 MRT node arithmetic shared by the dense and sparse kernels.
//
// Reference being replaced: Single_phase/LBM_3D_SinglePhase_Solver.py
//   lattice e[19], w[19], LR[19]            :85, :183-197
//   feq                                     :152-158
//   meq_vec                                 :209-215
//   colission (M, S_dig, Guo force, inv_M)  :222-241
//   streaming3 macroscopic part             :380-388
//
// Two evaluation modes, chosen per translation unit:
//   default      factored moment transform (pair sums/differences), FMA contraction on.
//                The inverse keeps the EFFECTIVE fp32 matrix of the reference: every entry
//                of np.linalg.inv(M) rounded to f32 (:83,:110) is a power of two times
//                t = f32(1/3) or n = f32(1/9) (or an exact power of two), so evaluating
//                the factored form with t and n reproduces the reference's systematic
//                rounding bias and only the (unbiased) order of roundings differs.
//   LBM_STRICT   the oracle's literal evaluation order (ascending-index sums, full
//                19x19 products with zeros skipped); the TU is compiled with -fmad=false,
//                which makes results bit-identical to oracle/ref_single_phase.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace d3q19 {

// X(s, ex, ey, ez, opposite)
#define D3Q19_DIRS(X)                                                                      \
    X(0, 0, 0, 0, 0)                                                                       \
    X(1, 1, 0, 0, 2) X(2, -1, 0, 0, 1) X(3, 0, 1, 0, 4) X(4, 0, -1, 0, 3)                  \
    X(5, 0, 0, 1, 6) X(6, 0, 0, -1, 5)                                                     \
    X(7, 1, 1, 0, 8) X(8, -1, -1, 0, 7) X(9, 1, -1, 0, 10) X(10, -1, 1, 0, 9)              \
    X(11, 1, 0, 1, 12) X(12, -1, 0, -1, 11) X(13, 1, 0, -1, 14) X(14, -1, 0, 1, 13)        \
    X(15, 0, 1, 1, 16) X(16, 0, -1, -1, 15) X(17, 0, 1, -1, 18) X(18, 0, -1, 1, 17)

// w[s] as stored in the reference's f32 field (:195-197)
#define W_REST ((float)(1.0 / 3.0))
#define W_AXIS ((float)(1.0 / 18.0))
#define W_DIAG ((float)(1.0 / 36.0))
__host__ __device__ constexpr float weight(int s) { return s == 0 ? W_REST : (s < 7 ? W_AXIS : W_DIAG); }

struct LbmParams {
    float S[19];          // S_dig :131
    float force[3];       // ext_f :134-136 (uniform force; a per-node force array replaces it)
    // Guo force term (:230-238).  The class divides its two parts by 3 and 9 (:236); the other
    // copy of the solver (Phase_change/LBM_3D_SinglePhase_Solver.py:235) does not.  guo_unscaled
    // selects the form (verification arithmetic); gc[] are the closed-form coefficients of the
    // production arithmetic: moment 0, 1, (3,5,7), (9,11), (13,14,15).
    int guo_unscaled;
    float gc[5];
    // form of a fixed-velocity face: 0 = the class, F = feq(1, u) for all 19 populations (:288);
    // 1 = the script copies (Single_phase/lbm_solver_3d.py:253), in place for s = 0..18:
    //     F[s] = feq(LR[s], 1, u) - F[LR[s]] + feq(s, 1, u)
    int vel_bc_script;
    int bc_type[6];       // x0,x1,y0,y1,z0,z1
    float bc_rho[6];
    float bc_vel[6][3];
};

// e_s . u with the zero components skipped (adding an exact zero changes nothing).
template <int EX, int EY, int EZ>
__device__ __forceinline__ float edotu(float ux, float uy, float uz) {
    float r = 0.f;
    bool first = true;
    if (EX != 0) { r = (EX > 0 ? ux : -ux); first = false; }
    if (EY != 0) { float t = (EY > 0 ? uy : -uy); r = first ? t : r + t; first = false; }
    if (EZ != 0) { float t = (EZ > 0 ? uz : -uz); r = first ? t : r + t; first = false; }
    return r;
}

// feq :152-158, same association as the source text
template <int S, int EX, int EY, int EZ>
__device__ __forceinline__ float feq(float rho, float ux, float uy, float uz) {
    const float eu = edotu<EX, EY, EZ>(ux, uy, uz);
    const float uv = ux * ux + uy * uy + uz * uz;
    return weight(S) * rho * (1.0f + 3.0f * eu + 4.5f * eu * eu - 1.5f * uv);
}

__device__ __forceinline__ void feq_all(float (&f)[19], float rho, float ux, float uy, float uz) {
#define X(s, ex, ey, ez, o) f[s] = feq<s, ex, ey, ez>(rho, ux, uy, uz);
    D3Q19_DIRS(X)
#undef X
}

#ifdef LBM_STRICT
// -------------------------------------------------------------------------------------------
// literal evaluation order (oracle/ref_single_phase.c)
// -------------------------------------------------------------------------------------------
static __constant__ float c_invM[361];   // inv_M :110, uploaded by the API (one copy per TU)

__device__ __forceinline__ void macro(const float (&f)[19], const float (&frc)[3], bool force,
                                      float &rho, float &ux, float &uy, float &uz) {
    float r = 0.f;
#pragma unroll
    for (int s = 0; s < 19; ++s) r = r + f[s];            // :380 .sum(), ascending
    float x = 0.f, y = 0.f, z = 0.f;
#define X(s, ex, ey, ez, o)                                                                  \
    if (ex != 0) x = x + (float)(ex) * f[s];                                                   \
    if (ey != 0) y = y + (float)(ey) * f[s];                                                   \
    if (ez != 0) z = z + (float)(ez) * f[s];
    D3Q19_DIRS(X)                                         // :382-383
#undef X
    x = x / r; y = y / r; z = z / r;                      // :387
    // :388   v += (f/2)/rho   (adding an exact zero when there is no force changes nothing)
    x = x + (frc[0] / 2.0f) / r;
    y = y + (frc[1] / 2.0f) / r;
    z = z + (frc[2] / 2.0f) / r;
    (void)force;
    rho = r; ux = x; uy = y; uz = z;
}

__device__ __forceinline__ void collide(float (&f)[19], const LbmParams &P, const float (&frc)[3], bool force,
                                        float rho, float ux, float uy, float uz) {
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
    float m[19], meq[19];
#pragma unroll
    for (int s = 0; s < 19; ++s) {                        // :226
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 19; ++l)
            if (M[s][l] != 0) acc = acc + (float)M[s][l] * f[l];
        m[s] = acc;
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) meq[s] = 0.f;            // :209-215
    meq[0] = rho; meq[3] = ux; meq[5] = uy; meq[7] = uz;
    meq[1] = ux * ux + uy * uy + uz * uz;
    meq[9] = 2.0f * ux * ux - uy * uy - uz * uz;
    meq[11] = uy * uy - uz * uz;
    meq[13] = ux * uy; meq[14] = uy * uz; meq[15] = ux * uz;
#pragma unroll
    for (int s = 0; s < 19; ++s) m[s] = m[s] - P.S[s] * (m[s] - meq[s]);   // :228
    if (force) {                                           // :230-238
        const float fx = frc[0], fy = frc[1], fz = frc[2];
#pragma unroll
        for (int s = 0; s < 19; ++s) {
            float f_guo = 0.f;
#pragma unroll
            for (int l = 0; l < 19; ++l) {
                if (M[s][l] == 0) continue;
                const float e0 = (float)EV[l][0], e1 = (float)EV[l][1], e2 = (float)EV[l][2];
                const float emv_f = (e0 - ux) * fx + (e1 - uy) * fy + (e2 - uz) * fz;
                const float ev = e0 * ux + e1 * uy + e2 * uz;
                const float ef = e0 * fx + e1 * fy + e2 * fz;
                const float term = P.guo_unscaled ? emv_f + (ev * ef) : emv_f / 3.0f + (ev * ef) / 9.0f;
                f_guo = f_guo + weight(l) * term * (float)M[s][l];
            }
            m[s] = m[s] + (1.0f - 0.5f * P.S[s]) * f_guo;
        }
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) {                        // :240-241
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
// -------------------------------------------------------------------------------------------
// production arithmetic: factored transforms
// -------------------------------------------------------------------------------------------
struct Moments { float m[19]; };

// m = M f  (:226) through pair sums / differences; exact algebra, ~70 adds.
__device__ __forceinline__ void forward(const float (&f)[19], float (&m)[19]) {
    const float px = f[1] + f[2], dx = f[1] - f[2];
    const float py = f[3] + f[4], dy = f[3] - f[4];
    const float pz = f[5] + f[6], dz = f[5] - f[6];
    const float a1 = f[7] + f[8], b1 = f[7] - f[8], a2 = f[9] + f[10], b2 = f[9] - f[10];
    const float c1 = f[11] + f[12], g1 = f[11] - f[12], c2 = f[13] + f[14], g2 = f[13] - f[14];
    const float h1 = f[15] + f[16], k1 = f[15] - f[16], h2 = f[17] + f[18], k2 = f[17] - f[18];
    const float Pxy = a1 + a2, Pxz = c1 + c2, Pyz = h1 + h2;
    const float A = b1 + b2, C = b1 - b2, B = g1 + g2, E = g1 - g2, D = k1 + k2, G = k1 - k2;
    const float P = px + py + pz, Q = Pxy + Pxz + Pyz;
    m[0] = f[0] + P + Q;
    m[1] = Q - f[0];
    m[2] = f[0] - 2.0f * P + Q;
    const float AB = A + B, CD = C + D, EG = E + G;
    m[3] = dx + AB; m[4] = AB - 2.0f * dx; m[16] = A - B;
    m[5] = dy + CD; m[6] = CD - 2.0f * dy; m[17] = D - C;
    m[7] = dz + EG; m[8] = EG - 2.0f * dz; m[18] = E - G;
    const float pyz = py + pz, t = (Pxy + Pxz) - 2.0f * Pyz;
    m[9] = 2.0f * px - pyz + t;
    m[10] = pyz - 2.0f * px + t;
    const float u = py - pz, w_ = Pxy - Pxz;
    m[11] = u + w_; m[12] = w_ - u;
    m[13] = a1 - a2; m[14] = h1 - h2; m[15] = c1 - c2;
}

// f = inv_M m (:240-241) with the reference's effective fp32 coefficients (see header).
__device__ __forceinline__ void inverse(const float (&m)[19], float (&f)[19]) {
    constexpr float t = (float)(1.0 / 3.0), n = (float)(1.0 / 9.0);
    constexpr float t2 = t * 0.5f, t4 = t * 0.25f, t8 = t * 0.125f, t16 = t * 0.0625f;
    constexpr float n2 = n * 0.5f, n4 = n * 0.25f, n8 = n * 0.125f;
    f[0] = t * m[0] - 0.5f * m[1] + t2 * m[2];
    const float cA = n2 * (m[0] - m[2]);
    const float d910 = m[9] - m[10], d1112 = m[11] - m[12];
    const float ax = cA + t4 * d910;
    const float ayz = cA - t8 * d910;
    const float ay = ayz + 0.125f * d1112, az = ayz - 0.125f * d1112;
    const float ox = t2 * (m[3] - m[4]), oy = t2 * (m[5] - m[6]), oz = t2 * (m[7] - m[8]);
    f[1] = ax + ox; f[2] = ax - ox;
    f[3] = ay + oy; f[4] = ay - oy;
    f[5] = az + oz; f[6] = az - oz;
    const float cD = n4 * m[0] + t8 * m[1] + n8 * m[2];
    const float s910 = m[9] + m[10], s1112 = m[11] + m[12];
    const float bx = cD + t16 * s910;
    const float bxy = bx + 0.0625f * s1112, bxz = bx - 0.0625f * s1112;
    const float byz = cD - t8 * s910;
    const float jx = t8 * (2.0f * m[3] + m[4]), jy = t8 * (2.0f * m[5] + m[6]),
                jz = t8 * (2.0f * m[7] + m[8]);
    {
        const float p = bxy + 0.25f * m[13], r = bxy - 0.25f * m[13];
        const float o1 = jx + jy + 0.125f * (m[16] - m[17]), o2 = jx - jy + 0.125f * (m[16] + m[17]);
        f[7] = p + o1; f[8] = p - o1; f[9] = r + o2; f[10] = r - o2;
    }
    {
        const float p = bxz + 0.25f * m[15], r = bxz - 0.25f * m[15];
        const float o1 = jx + jz + 0.125f * (m[18] - m[16]), o2 = jx - jz - 0.125f * (m[16] + m[18]);
        f[11] = p + o1; f[12] = p - o1; f[13] = r + o2; f[14] = r - o2;
    }
    {
        const float p = byz + 0.25f * m[14], r = byz - 0.25f * m[14];
        const float o1 = jy + jz + 0.125f * (m[17] - m[18]), o2 = jy - jz + 0.125f * (m[17] + m[18]);
        f[15] = p + o1; f[16] = p - o1; f[17] = r + o2; f[18] = r - o2;
    }
}

// rho, v of streaming3 (:380-388): rho = m0, momentum = (m3, m5, m7).
__device__ __forceinline__ void macro(const float (&f)[19], const float (&frc)[3], bool force,
                                      float &rho, float &ux, float &uy, float &uz) {
    const float px = f[1] + f[2], py = f[3] + f[4], pz = f[5] + f[6];
    const float a1 = f[7] + f[8], a2 = f[9] + f[10], c1 = f[11] + f[12], c2 = f[13] + f[14];
    const float h1 = f[15] + f[16], h2 = f[17] + f[18];
    const float r = f[0] + (px + py + pz) + ((a1 + a2) + (c1 + c2) + (h1 + h2));
    const float b1 = f[7] - f[8], b2 = f[9] - f[10], g1 = f[11] - f[12], g2 = f[13] - f[14];
    const float k1 = f[15] - f[16], k2 = f[17] - f[18];
    float x = (f[1] - f[2]) + ((b1 + b2) + (g1 + g2));
    float y = (f[3] - f[4]) + ((b1 - b2) + (k1 + k2));
    float z = (f[5] - f[6]) + ((g1 - g2) + (k1 - k2));
    const float inv = 1.0f / r;
    if (force) {
        x = (x + 0.5f * frc[0]) * inv;
        y = (y + 0.5f * frc[1]) * inv;
        z = (z + 0.5f * frc[2]) * inv;
    } else {
        x *= inv; y *= inv; z *= inv;
    }
    rho = r; ux = x; uy = y; uz = z;
}

// colission :222-241.  Guo term in closed form: sum_l w_l[((e_l-v).F)/3 + (e_l.v)(e_l.F)/9] M[s,l]
// is non-zero only for s in {0,1,3,5,7,9,11,13,14,15} (derived symbolically).
__device__ __forceinline__ void collide(float (&f)[19], const LbmParams &P, const float (&frc)[3], bool force,
                                        float rho, float ux, float uy, float uz) {
    float m[19];
    forward(f, m);
    const float uxx = ux * ux, uyy = uy * uy, uzz = uz * uz;
    // m -= S (m - meq), every row with its own rate (:228); rows with S = 0 stay exact.
    m[0] = m[0] - P.S[0] * (m[0] - rho);
    m[1] = m[1] - P.S[1] * (m[1] - (uxx + uyy + uzz));
    m[2] = m[2] - P.S[2] * m[2];
    m[3] = m[3] - P.S[3] * (m[3] - ux);
    m[4] = m[4] - P.S[4] * m[4];
    m[5] = m[5] - P.S[5] * (m[5] - uy);
    m[6] = m[6] - P.S[6] * m[6];
    m[7] = m[7] - P.S[7] * (m[7] - uz);
    m[8] = m[8] - P.S[8] * m[8];
    m[9] = m[9] - P.S[9] * (m[9] - (2.0f * uxx - uyy - uzz));
    m[10] = m[10] - P.S[10] * m[10];
    m[11] = m[11] - P.S[11] * (m[11] - (uyy - uzz));
    m[12] = m[12] - P.S[12] * m[12];
    m[13] = m[13] - P.S[13] * (m[13] - ux * uy);
    m[14] = m[14] - P.S[14] * (m[14] - uy * uz);
    m[15] = m[15] - P.S[15] * (m[15] - ux * uz);
    m[16] = m[16] - P.S[16] * m[16];
    m[17] = m[17] - P.S[17] * m[17];
    m[18] = m[18] - P.S[18] * m[18];
    if (force) {                                                               // :230-238
        const float fx = frc[0], fy = frc[1], fz = frc[2];
        const float xx = fx * ux, yy = fy * uy, zz = fz * uz;
        const float vf = xx + yy + zz;
        // class form: gc = (-8/27, 2/81, 1/9, 2/81, 1/81); un-scaled: (-2/3, 2/9, 1/3, 2/9, 1/9)
        m[0] += (1.0f - 0.5f * P.S[0]) * P.gc[0] * vf;
        m[1] += (1.0f - 0.5f * P.S[1]) * P.gc[1] * vf;
        m[3] += (1.0f - 0.5f * P.S[3]) * P.gc[2] * fx;
        m[5] += (1.0f - 0.5f * P.S[5]) * P.gc[2] * fy;
        m[7] += (1.0f - 0.5f * P.S[7]) * P.gc[2] * fz;
        m[9] += (1.0f - 0.5f * P.S[9]) * P.gc[3] * (2.0f * xx - yy - zz);
        m[11] += (1.0f - 0.5f * P.S[11]) * P.gc[3] * (yy - zz);
        m[13] += (1.0f - 0.5f * P.S[13]) * P.gc[4] * (fx * uy + fy * ux);
        m[14] += (1.0f - 0.5f * P.S[14]) * P.gc[4] * (fy * uz + fz * uy);
        m[15] += (1.0f - 0.5f * P.S[15]) * P.gc[4] * (fx * uz + fz * ux);
    }
    inverse(m, f);
}
#endif  // LBM_STRICT

}  // namespace d3q19
