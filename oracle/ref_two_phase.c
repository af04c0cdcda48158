/* CPU restatement (oracle) of the reference two-phase colour-gradient step, in C.
 *
 * TEST INFRASTRUCTURE ONLY; parity pin: the reference script run through tests/taichi_shim (bit for
 * bit up to the order of its float atomics, tests/test_reference_pin.py) -- see oracle/ref_two_phase.py, whose operations
 * this file repeats one for one (the -ffp-contract=off build is bit-identical to the NumPy
 * form; tests/test_oracle_two_phase.py).  Line numbers cite
 * /root/reference/2phase/lbm_solver_3d_2phase.py.
 *
 * The colour push of colission (:365-372, float atomics in a parallel loop, order undefined
 * in the reference) is evaluated in pull form, ascending direction at the destination.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#define SUF f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)
#define R(x) ((REAL)(x))

static const int Mi[19][19] = {
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
static const int LRi[19] = {0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17};
static const int Ei[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
    {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1}, {1, 0, -1},
    {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};

typedef struct {
    int nx, ny, nz;
    int bc_type[6];       /* flow faces: 0 periodic, 1 pressure, 2 velocity */
    int bc_psi_type[6];   /* phase-field faces: 0 periodic, 1 constant */
    REAL invM[19 * 19];
    REAL w[19];
    REAL force[3];
    REAL bc_rho[6];
    REAL bc_psi_val[6];
    REAL psi_solid, CapA;
    REAL wl, wg, lg0, l1, l2, g1, g2;   /* :100-108 */
} FN(ref2p_params);
typedef FN(ref2p_params) params_t;

static inline size_t nidx(const params_t *p, int i, int j, int k) {
    return ((size_t)i * p->ny + j) * p->nz + k;
}
static inline int wrap(int i, int n) { return i < 0 ? n - 1 : (i > n - 1 ? 0 : i); }
/* periodic_index_for_psi :390-428 */
static inline int wrap_psi(int i, int n, int type_lo, int type_hi) {
    if (i < 0) return type_lo == 0 ? n - 1 : 0;
    if (i > n - 1) return type_hi == 0 ? 0 : n - 1;
    return i;
}
static inline REAL feq(const params_t *p, int k, REAL rho, const REAL *u) {
    REAL eu = R(Ei[k][0]) * u[0] + R(Ei[k][1]) * u[1] + R(Ei[k][2]) * u[2];
    REAL uv = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    return p->w[k] * rho * (R(1.0) + R(3.0) * eu + R(4.5) * eu * eu - R(1.5) * uv);
}

/* colission :302-363 without the push: writes f, g_r, g_b ([N][19]) */
void FN(ref2p_collide)(const params_t *p, const int8_t *solid, const REAL *F, const REAL *rho,
                       const REAL *v, const REAL *psi, const REAL *rho_r, const REAL *rho_b, REAL *f,
                       REAL *g_r, REAL *g_b) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < p->nx; ++i)
        for (int j = 0; j < p->ny; ++j)
            for (int k = 0; k < p->nz; ++k) {
                const size_t c = nidx(p, i, j, k);
                if (solid[c] != 0) continue;
                /* Compute_C :259-275 */
                REAL C[3] = {R(0), R(0), R(0)};
                int ind_S = 0;
                for (int s = 0; s < 19; ++s) {
                    size_t ip = nidx(p, wrap_psi(i + Ei[s][0], p->nx, p->bc_psi_type[0], p->bc_psi_type[1]),
                                     wrap_psi(j + Ei[s][1], p->ny, p->bc_psi_type[2], p->bc_psi_type[3]),
                                     wrap_psi(k + Ei[s][2], p->nz, p->bc_psi_type[4], p->bc_psi_type[5]));
                    REAL val;
                    if (solid[ip] == 0) val = psi[ip];
                    else { ind_S = 1; val = p->psi_solid; }
                    for (int d = 0; d < 3; ++d) C[d] = C[d] + R(3.0) * p->w[s] * R(Ei[s][d]) * val;
                }
                REAL dlt = rho_r[c] - rho_b[c];
                if (dlt < 0) dlt = -dlt;
                if (dlt > R(0.9) && ind_S == 1) { C[0] = R(0); C[1] = R(0); C[2] = R(0); }
                const REAL cc = (REAL)sqrt((double)(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]));
                REAL nrm[3] = {R(0), R(0), R(0)};
                if (cc > 0) { nrm[0] = C[0] / cc; nrm[1] = C[1] / cc; nrm[2] = C[2] / cc; }
                const REAL *u = v + c * 3;
                const REAL *Fc = F + c * 19;
                REAL m[19], meq[19];
                for (int s = 0; s < 19; ++s) {
                    REAL acc = R(0);
                    for (int l = 0; l < 19; ++l)
                        if (Mi[s][l] != 0) acc = acc + R(Mi[s][l]) * Fc[l];
                    m[s] = acc;
                }
                for (int s = 0; s < 19; ++s) meq[s] = R(0);
                meq[0] = rho[c];
                meq[3] = u[0]; meq[5] = u[1]; meq[7] = u[2];
                meq[1] = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
                meq[9] = R(2) * u[0] * u[0] - u[1] * u[1] - u[2] * u[2];
                meq[11] = u[1] * u[1] - u[2] * u[2];
                meq[13] = u[0] * u[1]; meq[14] = u[1] * u[2]; meq[15] = u[0] * u[2];
                meq[1] = meq[1] + p->CapA * cc;
                meq[9] = meq[9] + R(0.5) * p->CapA * cc * (R(2) * nrm[0] * nrm[0] - nrm[1] * nrm[1] - nrm[2] * nrm[2]);
                meq[11] = meq[11] + R(0.5) * p->CapA * cc * (nrm[1] * nrm[1] - nrm[2] * nrm[2]);
                meq[13] = meq[13] + R(0.5) * p->CapA * cc * (nrm[0] * nrm[1]);
                meq[14] = meq[14] + R(0.5) * p->CapA * cc * (nrm[1] * nrm[2]);
                meq[15] = meq[15] + R(0.5) * p->CapA * cc * (nrm[0] * nrm[2]);
                /* Compute_S_local :278-299 */
                const REAL ps = psi[c];
                REAL sv;
                if (ps > 0) sv = ps > R(0.1) ? p->wl : p->lg0 + p->l1 * ps + p->l2 * ps * ps;
                else sv = ps < R(-0.1) ? p->wg : p->lg0 + p->g1 * ps + p->g2 * ps * ps;
                const REAL so = R(8.0) * (R(2.0) - sv) / (R(8.0) - sv);
                const REAL S[19] = {R(0), sv, sv, R(0), so, R(0), so, R(0), so, sv, sv, sv, sv, sv, sv, sv, so, so, so};
                const REAL *fo = p->force;
                for (int s = 0; s < 19; ++s) {
                    m[s] = m[s] - S[s] * (m[s] - meq[s]);
                    REAL guo = R(0);
                    for (int l = 0; l < 19; ++l) {
                        if (Mi[s][l] == 0) continue;
                        REAL e0 = R(Ei[l][0]), e1 = R(Ei[l][1]), e2 = R(Ei[l][2]);
                        REAL emu_f = (e0 - u[0]) * fo[0] + (e1 - u[1]) * fo[1] + (e2 - u[2]) * fo[2];
                        REAL eu = e0 * u[0] + e1 * u[1] + e2 * u[2];
                        REAL ef = e0 * fo[0] + e1 * fo[1] + e2 * fo[2];
                        guo = guo + p->w[l] * (emu_f + (eu * ef)) * R(Mi[s][l]);
                    }
                    m[s] = m[s] + (R(1) - R(0.5) * S[s]) * guo;
                }
                REAL *fc = f + c * 19, *gr = g_r + c * 19, *gb = g_b + c * 19;
                for (int s = 0; s < 19; ++s) {
                    REAL acc = R(0);
                    for (int l = 0; l < 19; ++l) {
                        REAL a = p->invM[s * 19 + l];
                        if (a != R(0)) acc = acc + a * m[l];
                    }
                    fc[s] = acc;
                    gr[s] = feq(p, s, rho_r[c], u);
                    gb[s] = feq(p, s, rho_b[c], u);
                }
                if (cc > 0) {
                    for (int kk = 1; kk < 19; kk += 2) {
                        REAL ef = R(Ei[kk][0]) * C[0] + R(Ei[kk][1]) * C[1] + R(Ei[kk][2]) * C[2];
                        REAL cs = gr[kk] < gr[kk + 1] ? gr[kk] : gr[kk + 1];
                        cs = cs < gb[kk] ? cs : gb[kk];
                        cs = cs < gb[kk + 1] ? cs : gb[kk + 1];
                        cs = cs * (ef / cc);
                        gr[kk] = gr[kk] + cs;
                        gr[kk + 1] = gr[kk + 1] - cs;
                        gb[kk] = gb[kk] - cs;
                        gb[kk + 1] = gb[kk + 1] + cs;
                    }
                }
            }
}

/* :365-372 in pull form, ascending s at the destination; adds into rhor / rhob */
void FN(ref2p_accumulate)(const params_t *p, const int8_t *solid, const REAL *g_r, const REAL *g_b,
                          REAL *rhor, REAL *rhob) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < p->nx; ++i)
        for (int j = 0; j < p->ny; ++j)
            for (int k = 0; k < p->nz; ++k) {
                const size_t c = nidx(p, i, j, k);
                if (solid[c] != 0) continue;
                REAL ar = rhor[c], ab = rhob[c];
                for (int s = 0; s < 19; ++s) {
                    size_t src = nidx(p, wrap(i - Ei[s][0], p->nx), wrap(j - Ei[s][1], p->ny), wrap(k - Ei[s][2], p->nz));
                    if (solid[src] == 0) { ar = ar + g_r[src * 19 + s]; ab = ab + g_b[src * 19 + s]; }
                    else { ar = ar + g_r[c * 19 + LRi[s]]; ab = ab + g_b[c * 19 + LRi[s]]; }
                }
                rhor[c] = ar;
                rhob[c] = ab;
            }
}

/* :431-442 */
void FN(ref2p_streaming1)(const params_t *p, const int8_t *solid, const REAL *f, REAL *F) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < p->nx; ++i)
        for (int j = 0; j < p->ny; ++j)
            for (int k = 0; k < p->nz; ++k) {
                size_t c = nidx(p, i, j, k);
                if (solid[c] != 0) continue;
                for (int s = 0; s < 19; ++s) {
                    size_t ip = nidx(p, wrap(i + Ei[s][0], p->nx), wrap(j + Ei[s][1], p->ny), wrap(k + Ei[s][2], p->nz));
                    if (solid[ip] == 0) F[ip * 19 + s] = f[c * 19 + s];
                    else F[c * 19 + LRi[s]] = f[c * 19 + s];
                }
            }
}

/* :491-583 */
void FN(ref2p_boundary_condition)(const params_t *p, const int8_t *solid, const REAL *v, REAL *F) {
    const int n[3] = {p->nx, p->ny, p->nz};
    const REAL zero_u[3] = {R(0), R(0), R(0)};   /* bc_vel_* fields are never written */
    for (int face = 0; face < 6; ++face) {
        int t = p->bc_type[face];
        if (t == 0) continue;
        int axis = face / 2, side = face % 2;
        int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        int pos = side == 0 ? 0 : n[axis] - 1;
        int pin = side == 0 ? 1 : n[axis] - 2;
#pragma omp parallel for collapse(2) schedule(static)
        for (int q = 0; q < n[a1]; ++q)
            for (int r = 0; r < n[a2]; ++r) {
                int ijk[3], ijk_in[3];
                ijk[axis] = pos; ijk[a1] = q; ijk[a2] = r;
                ijk_in[axis] = pin; ijk_in[a1] = q; ijk_in[a2] = r;
                size_t c = nidx(p, ijk[0], ijk[1], ijk[2]);
                if (solid[c] != 0) continue;
                if (t == 1) {
                    size_t cin = nidx(p, ijk_in[0], ijk_in[1], ijk_in[2]);
                    const REAL *u = solid[cin] > 0 ? v + cin * 3 : v + c * 3;
                    for (int s = 0; s < 19; ++s) F[c * 19 + s] = feq(p, s, p->bc_rho[face], u);
                } else {
                    for (int s = 0; s < 19; ++s)
                        F[c * 19 + s] = feq(p, LRi[s], R(1.0), zero_u) - F[c * 19 + LRi[s]] + feq(p, s, R(1.0), zero_u);
                }
            }
    }
}

/* :587-605 */
void FN(ref2p_streaming3)(const params_t *p, const int8_t *solid, const REAL *F, REAL *f, REAL *rho, REAL *v,
                          REAL *psi, REAL *rho_r, REAL *rho_b, REAL *rhor, REAL *rhob) {
    const size_t n = (size_t)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        if (solid[c] != 0) continue;
        rho_r[c] = rhor[c];
        rho_b[c] = rhob[c];
        rhor[c] = R(0);
        rhob[c] = R(0);
        REAL r = R(0), u[3] = {R(0), R(0), R(0)};
        for (int s = 0; s < 19; ++s) {
            f[c * 19 + s] = F[c * 19 + s];
            r = r + F[c * 19 + s];
            for (int d = 0; d < 3; ++d)
                if (Ei[s][d] != 0) u[d] = u[d] + R(Ei[s][d]) * F[c * 19 + s];
        }
        for (int d = 0; d < 3; ++d) {
            u[d] = u[d] / r;
            u[d] = u[d] + (p->force[d] / R(2)) / r;
        }
        rho[c] = r;
        v[c * 3] = u[0]; v[c * 3 + 1] = u[1]; v[c * 3 + 2] = u[2];
        psi[c] = rho_r[c] - rho_b[c] / (rho_r[c] + rho_b[c]);
    }
}

/* :445-486 */
void FN(ref2p_boundary_condition_psi)(const params_t *p, const int8_t *solid, REAL *psi, REAL *rho_r, REAL *rho_b) {
    const int n[3] = {p->nx, p->ny, p->nz};
    for (int face = 0; face < 6; ++face) {
        if (p->bc_psi_type[face] != 1) continue;
        int axis = face / 2, side = face % 2;
        int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        int pos = side == 0 ? 0 : n[axis] - 1;
        for (int q = 0; q < n[a1]; ++q)
            for (int r = 0; r < n[a2]; ++r) {
                int ijk[3];
                ijk[axis] = pos; ijk[a1] = q; ijk[a2] = r;
                size_t c = nidx(p, ijk[0], ijk[1], ijk[2]);
                if (solid[c] != 0) continue;
                psi[c] = p->bc_psi_val[face];
                rho_r[c] = (p->bc_psi_val[face] + R(1.0)) / R(2.0);
                rho_b[c] = R(1.0) - rho_r[c];
            }
    }
}

/* main loop body :626-632, nsteps times; g_r, g_b are scratch [N][19] */
void FN(ref2p_step)(const params_t *p, const int8_t *solid, REAL *f, REAL *F, REAL *rho, REAL *v, REAL *psi,
                    REAL *rho_r, REAL *rho_b, REAL *rhor, REAL *rhob, REAL *g_r, REAL *g_b, int nsteps) {
    for (int it = 0; it < nsteps; ++it) {
        FN(ref2p_collide)(p, solid, F, rho, v, psi, rho_r, rho_b, f, g_r, g_b);
        FN(ref2p_accumulate)(p, solid, g_r, g_b, rhor, rhob);
        FN(ref2p_streaming1)(p, solid, f, F);
        FN(ref2p_boundary_condition)(p, solid, v, F);
        FN(ref2p_streaming3)(p, solid, F, f, rho, v, psi, rho_r, rho_b, rhor, rhob);
        FN(ref2p_boundary_condition_psi)(p, solid, psi, rho_r, rho_b);
    }
}

size_t FN(ref2p_sizeof_params)(void) { return sizeof(params_t); }
