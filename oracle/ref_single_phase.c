/* CPU restatement (oracle) of the reference single-phase D3Q19 MRT time step, in C.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (parity checker), __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never linked into, imported by
 * or called from the product library (taichi_lbm3d_b200/csrc).
 *
 * PARITY PIN: the reference holds no golden vectors and Taichi cannot run in this image;
 * the pin is the reference's own source executed through tests/taichi_shim, reproduced bit
 * for bit (tests/test_reference_pin.py; full statement in oracle/ref_single_phase.py).  This file is the
 * same algorithm as that NumPy form, operation for operation (built with
 * -ffp-contract=off the two are bit-identical; tests/test_oracle.py checks it), kept
 * in the reference's own four-pass AoS structure so that, threaded with OpenMP, it
 * stands in for the reference's ti.init(arch=ti.cpu) path as the CPU baseline.
 *
 * Line numbers cite /root/reference/Single_phase/LBM_3D_SinglePhase_Solver.py.
 *
 * Compile twice: -DREAL=float -DSUF=f32 and -DREAL=double -DSUF=f64 (oracle/Makefile).
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#ifndef REAL
#define REAL float
#define SUF f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)
#define R(x) ((REAL)(x))

/* :90-108 */
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
/* :85 */
static const int LRi[19] = {0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17};
/* :183-187 */
static const int Ei[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
    {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1}, {1, 0, -1},
    {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};

/* Runtime parameters; the Python side fills this (ctypes) exactly as
 * init_simulation :118-149 computes them (float64 math, one rounding to REAL). */
typedef struct {
    int nx, ny, nz;
    int force_flag;          /* :137-140 */
    int bc_type[6];          /* x0,x1,y0,y1,z0,z1 : 0 periodic, 1 pressure, 2 velocity */
    REAL S[19];              /* :131 */
    REAL invM[19 * 19];      /* :110 */
    REAL w[19];              /* :195-197 */
    REAL force[3];           /* :134-136 */
    REAL bc_rho[6];
    REAL bc_vel[6][3];
    const REAL *force_field; /* [n][3] per-node force = cal_local_force(i,j,k) :217-220, or NULL */
    int guo_unscaled;        /* 1: Phase_change/LBM_3D_SinglePhase_Solver.py:235 (no /3, /9) */
    int vel_bc_script;       /* 1: Single_phase/lbm_solver_3d.py:253 (in-place velocity form) */
    const REAL *ns;          /* [n] solid fraction per node (Grey_Scale/lbm_solver_3d_Macro_Sukop.py:42), or NULL */
} FN(ref_params);
typedef FN(ref_params) params_t;

static inline size_t nidx(const params_t *p, int i, int j, int k) {
    return ((size_t)i * p->ny + j) * p->nz + k;
}

/* :152-158 */
static inline REAL feq(const params_t *p, int k, REAL rho, const REAL *u) {
    REAL eu = R(Ei[k][0]) * u[0] + R(Ei[k][1]) * u[1] + R(Ei[k][2]) * u[2];
    REAL uv = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    return p->w[k] * rho * (R(1.0) + R(3.0) * eu + R(4.5) * eu * eu - R(1.5) * uv);
}

/* :222-241 */
void FN(ref_sp_colission)(const params_t *p, const int8_t *solid, const REAL *F, const REAL *rho,
                          const REAL *v, REAL *f) {
    const size_t n = (size_t)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        if (solid[c] != 0) continue;
        const REAL *Fc = F + c * 19;
        const REAL *u = v + c * 3;
        REAL m[19], meq[19];
        for (int s = 0; s < 19; ++s) {                      /* :226 */
            REAL acc = R(0);
            for (int l = 0; l < 19; ++l)
                if (Mi[s][l] != 0) acc = acc + R(Mi[s][l]) * Fc[l];
            m[s] = acc;
        }
        for (int s = 0; s < 19; ++s) meq[s] = R(0);         /* :209-215 */
        meq[0] = rho[c];
        meq[3] = u[0]; meq[5] = u[1]; meq[7] = u[2];
        meq[1] = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        meq[9] = R(2) * u[0] * u[0] - u[1] * u[1] - u[2] * u[2];
        meq[11] = u[1] * u[1] - u[2] * u[2];
        meq[13] = u[0] * u[1]; meq[14] = u[1] * u[2]; meq[15] = u[0] * u[2];
        for (int s = 0; s < 19; ++s) m[s] = m[s] - p->S[s] * (m[s] - meq[s]);   /* :228 */
        if (p->force_flag == 1) {                            /* :230-238 */
            const REAL *fo = p->force_field ? p->force_field + c * 3 : p->force;      /* :231 */
            for (int s = 0; s < 19; ++s) {
                REAL f_guo = R(0);
                for (int l = 0; l < 19; ++l) {
                    if (Mi[s][l] == 0) continue;
                    REAL e0 = R(Ei[l][0]), e1 = R(Ei[l][1]), e2 = R(Ei[l][2]);
                    REAL emv_f = (e0 - u[0]) * fo[0] + (e1 - u[1]) * fo[1] + (e2 - u[2]) * fo[2];
                    REAL ev = e0 * u[0] + e1 * u[1] + e2 * u[2];
                    REAL ef = e0 * fo[0] + e1 * fo[1] + e2 * fo[2];
                    REAL term = p->guo_unscaled ? emv_f + (ev * ef) : emv_f / R(3.0) + (ev * ef) / R(9.0);
                    f_guo = f_guo + p->w[l] * term * R(Mi[s][l]);
                }
                m[s] = m[s] + (R(1) - R(0.5) * p->S[s]) * f_guo;
            }
        }
        REAL *fc = f + c * 19;
        for (int s = 0; s < 19; ++s) {                      /* :240-241 */
            REAL acc = R(0);
            for (int l = 0; l < 19; ++l) {
                REAL a = p->invM[s * 19 + l];
                if (a != R(0)) acc = acc + a * m[l];
            }
            fc[s] = acc;
        }
    }
}

/* :247-257 */
static inline int wrap(int i, int n) { return i < 0 ? n - 1 : (i > n - 1 ? 0 : i); }

/* :259-268  push.  Every F slot of a fluid node is written exactly once, so the
 * parallel loop is race-free. */
void FN(ref_sp_streaming1)(const params_t *p, const int8_t *solid, const REAL *f, REAL *F) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < p->nx; ++i)
        for (int j = 0; j < p->ny; ++j)
            for (int k = 0; k < p->nz; ++k) {
                size_t c = nidx(p, i, j, k);
                if (solid[c] != 0) continue;
                for (int s = 0; s < 19; ++s) {
                    size_t ip = nidx(p, wrap(i + Ei[s][0], p->nx), wrap(j + Ei[s][1], p->ny),
                                     wrap(k + Ei[s][2], p->nz));
                    if (solid[ip] == 0) F[ip * 19 + s] = f[c * 19 + s];
                    else F[c * 19 + LRi[s]] = f[c * 19 + s];
                }
            }
}

/* Grey_Scale/lbm_solver_3d_Macro_Sukop.py:233-247 (streaming0 + streaming1): blend with the opposite
 * population of the node ahead, weighted by the node's solid fraction, then push to ALL neighbours
 * (no bounce-back; solid nodes never push, so the links that leave them keep their old value).
 * f2 is evaluated on the fly: f is not written between the two passes. */
void FN(ref_sp_streaming_grey)(const params_t *p, const int8_t *solid, const REAL *f, REAL *F) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < p->nx; ++i)
        for (int j = 0; j < p->ny; ++j)
            for (int k = 0; k < p->nz; ++k) {
                size_t c = nidx(p, i, j, k);
                if (solid[c] != 0) continue;
                for (int s = 0; s < 19; ++s) {
                    size_t ip = nidx(p, wrap(i + Ei[s][0], p->nx), wrap(j + Ei[s][1], p->ny),
                                     wrap(k + Ei[s][2], p->nz));
                    REAL d = f[ip * 19 + LRi[s]] - f[c * 19 + s];
                    REAL t = p->ns[c] * d;
                    F[ip * 19 + s] = f[c * 19 + s] + t;
                }
            }
}

/* :272-370  faces x0,x1,y0,y1,z0,z1 in order, each a complete loop before the next. */
void FN(ref_sp_boundary_condition)(const params_t *p, const int8_t *solid, const REAL *v, REAL *F) {
    const int n[3] = {p->nx, p->ny, p->nz};
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
                } else if (p->vel_bc_script) {
                    /* Single_phase/lbm_solver_3d.py:253, in place for s = 0..18 */
                    for (int s = 0; s < 19; ++s)
                        F[c * 19 + s] = feq(p, LRi[s], R(1.0), p->bc_vel[face]) - F[c * 19 + LRi[s]] +
                                        feq(p, s, R(1.0), p->bc_vel[face]);
                } else {
                    for (int s = 0; s < 19; ++s) F[c * 19 + s] = feq(p, s, R(1.0), p->bc_vel[face]);
                }
            }
    }
}

/* :372-392 */
void FN(ref_sp_streaming3)(const params_t *p, const int8_t *solid, const REAL *F, REAL *f, REAL *rho,
                           REAL *v) {
    const size_t n = (size_t)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        if (solid[c] == 0) {
            REAL r = R(0), u[3] = {R(0), R(0), R(0)};
            for (int s = 0; s < 19; ++s) f[c * 19 + s] = F[c * 19 + s];
            for (int s = 0; s < 19; ++s) r = r + F[c * 19 + s];
            for (int s = 0; s < 19; ++s)
                for (int d = 0; d < 3; ++d)
                    if (Ei[s][d] != 0) u[d] = u[d] + R(Ei[s][d]) * F[c * 19 + s];
            const REAL *fo = p->force_field ? p->force_field + c * 3 : p->force;      /* :385 */
            for (int d = 0; d < 3; ++d) {
                u[d] = u[d] / r;
                u[d] = u[d] + (fo[d] / R(2)) / r;
            }
            rho[c] = r;
            v[c * 3 + 0] = u[0]; v[c * 3 + 1] = u[1]; v[c * 3 + 2] = u[2];
        } else {
            rho[c] = R(1.0);
            v[c * 3 + 0] = R(0); v[c * 3 + 1] = R(0); v[c * 3 + 2] = R(0);
        }
    }
}

/* :477-481, repeated nsteps times */
void FN(ref_sp_step)(const params_t *p, const int8_t *solid, REAL *f, REAL *F, REAL *rho, REAL *v,
                     int nsteps) {
    for (int it = 0; it < nsteps; ++it) {
        FN(ref_sp_colission)(p, solid, F, rho, v, f);
        if (p->ns) FN(ref_sp_streaming_grey)(p, solid, f, F);
        else FN(ref_sp_streaming1)(p, solid, f, F);
        FN(ref_sp_boundary_condition)(p, solid, v, F);
        FN(ref_sp_streaming3)(p, solid, F, f, rho, v);
    }
}

/* :394-402 */
REAL FN(ref_sp_max_v)(const params_t *p, const REAL *v) {
    const size_t n = (size_t)p->nx * p->ny * p->nz;
    REAL best = R(-1e10);
#pragma omp parallel for reduction(max : best) schedule(static)
    for (size_t c = 0; c < n; ++c) {
        REAL nr = (REAL)sqrt((double)(v[c * 3] * v[c * 3] + v[c * 3 + 1] * v[c * 3 + 1] + v[c * 3 + 2] * v[c * 3 + 2]));
        if (nr > best) best = nr;
    }
    return best;
}

size_t FN(ref_sp_sizeof_params)(void) { return sizeof(params_t); }
