// C-ABI layer of the two-phase colour-gradient solver (include/lbm3d_2phase.h).
// Reference: 2phase/lbm_solver_3d_2phase.py (line numbers below).  Dense storage, one GPU.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <thrust/iterator/transform_iterator.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/lbm3d_2phase.h"
#include "lbm2p_kernels.cuh"
#include "lbm_geometry.cuh"

namespace {
thread_local std::string g2_create_error;
struct Face2 {
    int type = 0;
    float rho = 1.0f;
    float vel[3] = {0.f, 0.f, 0.f};
};
}  // namespace

struct lbm2p_ctx {
    lbm2p_config cfg{};
    std::string err;
    int block = 256;
    // parameters (script defaults :18-39)
    double niu_l = 0.1, niu_g = 0.1, psi_solid = 0.7, CapA = 0.005;
    float force[3] = {5.0e-5f, -2e-5f, 0.0f};
    Face2 face[6];
    int bc_psi_type[6] = {1, 0, 0, 0, 0, 0};
    float bc_psi_val[6] = {-1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
    float invM[361]{};
    bool inited = false;
    size_t N = 0, nf = 0;
    uint32_t nzp = 0, prow = 0;
    size_t fsize = 0, pad = 0, npad = 0;
    int spec = 1;
    // device
    int8_t *d_solid = nullptr;
    float *d_psi0 = nullptr;          // input phase field (kept for the solid nodes of get_psi)
    uint32_t *d_flags = nullptr;
    uint8_t *d_cls = nullptr;
    float *d_fbase[2] = {nullptr, nullptr}, *d_f[2] = {nullptr, nullptr};
    float4 *d_recA = nullptr, *d_recC = nullptr;
    float2 *d_recB = nullptr;
    float *d_psibase = nullptr, *d_psi = nullptr;
    float *d_rho_r = nullptr, *d_rho_b = nullptr;
    float *d_rho = nullptr, *d_v = nullptr, *d_F = nullptr;
    float *d_vbc = nullptr;
    uint32_t vbc_off[6]{};
    float *d_scalar = nullptr;
    // state machine
    int cur = 0;
    bool pipe_valid = false;      // d_f[cur] + records hold the collision of the pending step
    bool colour_valid = true;     // rho_r, rho_b, psi are those of the last completed step
    bool macro_valid = true, F_valid = true;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
};

#define CTX2(ctx)                                                                              \
    if ((ctx) == nullptr) return -1;
#define FAIL2(ctx, code, ...)                                                                  \
    do {                                                                                       \
        char _b[512];                                                                          \
        snprintf(_b, sizeof _b, __VA_ARGS__);                                                  \
        (ctx)->err = _b;                                                                       \
        return (code);                                                                         \
    } while (0)
#define CU2(ctx, call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _b[512];                                                                      \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                                      \
            (ctx)->err = _b;                                                                   \
            return _e == cudaErrorMemoryAllocation ? -3 : -2;                                  \
        }                                                                                      \
    } while (0)

namespace {

// init :173-186 on the node-linear arrays; the working psi array holds psi_solid at solid nodes
__global__ void k2p_init_state(const int8_t *__restrict__ solid, const float *__restrict__ psi0, size_t n,
                               float psi_solid, float *psi, float *rho_r, float *rho_b, float *rho, float *v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (solid[i] == 0) {
        const float p = psi0[i];
        psi[i] = p;
        const float rr = (p + 1.0f) / 2.0f;
        rho_r[i] = rr;
        rho_b[i] = 1.0f - rr;
        rho[i] = 1.0f;
    } else {
        psi[i] = psi_solid;
        rho_r[i] = 0.f;
        rho_b[i] = 0.f;
        rho[i] = 0.f;
    }
    v[3 * i] = 0.f; v[3 * i + 1] = 0.f; v[3 * i + 2] = 0.f;
}

__global__ void k2p_fill_F(const int8_t *__restrict__ solid, float *F, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 19) F[i] = solid[i / 19] == 0 ? d3q19::weight((int)(i % 19)) : 0.f;
}

__global__ void k2p_merge_psi(const int8_t *__restrict__ solid, const float *__restrict__ psi0,
                              const float *__restrict__ psi, float *out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = solid[i] == 0 ? psi[i] : psi0[i];
}

__global__ void k2p_bake_psi(const int8_t *__restrict__ solid, float *psi, float psi_solid, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && solid[i] != 0) psi[i] = psi_solid;
}

void free2(lbm2p_ctx *c) {
    cudaFree(c->d_flags); cudaFree(c->d_cls); cudaFree(c->d_fbase[0]); cudaFree(c->d_fbase[1]);
    cudaFree(c->d_recA); cudaFree(c->d_recB); cudaFree(c->d_recC); cudaFree(c->d_psibase); cudaFree(c->d_rho_r); cudaFree(c->d_rho_b);
    cudaFree(c->d_rho); cudaFree(c->d_v); cudaFree(c->d_F); cudaFree(c->d_vbc); cudaFree(c->d_scalar);
    c->d_flags = nullptr; c->d_cls = nullptr; c->d_fbase[0] = c->d_fbase[1] = nullptr;
    c->d_recA = c->d_recC = nullptr; c->d_recB = nullptr; c->d_psibase = nullptr; c->d_rho_r = c->d_rho_b = nullptr;
    c->d_rho = c->d_v = c->d_F = nullptr; c->d_vbc = nullptr; c->d_scalar = nullptr;
}

void fill2(const lbm2p_ctx *c, Step2Args &A, const float *fin, float *fout) {
    memset(&A, 0, sizeof A);
    StepArgs &a = A.a;
    static const int e[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
        {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1},
        {1, 0, -1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};
    const long long sy = (long long)c->prow, sx = (long long)c->cfg.ny * sy;
    for (int s = 0; s < 19; ++s) {
        a.pown[s] = fin ? fin + (size_t)s * c->nzp : nullptr;
        a.ppull[s] = fin ? a.pown[s] - (e[s][0] * sx + e[s][1] * sy + e[s][2]) : nullptr;
        a.pout[s] = fout ? fout + (size_t)s * c->nzp : nullptr;
    }
    a.stride = 0;
    a.row_first = 0;
    a.row_count = (uint32_t)(c->cfg.nx * c->cfg.ny);
    a.prow = c->prow;
    a.spec = c->spec;
    a.nx = c->cfg.nx; a.ny = c->cfg.ny; a.nz = c->cfg.nz;
    a.flags = c->d_flags;
    a.cls = c->d_cls;
    a.rho = c->d_rho; a.v = c->d_v; a.F = nullptr;
    a.vbc = c->d_vbc;
    for (int i = 0; i < 6; ++i) a.vbc_off[i] = c->vbc_off[i];
    a.force = (fabsf(c->force[0]) > 0.f || fabsf(c->force[1]) > 0.f || fabsf(c->force[2]) > 0.f) ? 1 : 0;
    for (int i = 0; i < 3; ++i) a.P.force[i] = c->force[i];
    for (int i = 0; i < 6; ++i) {
        a.P.bc_type[i] = c->face[i].type;
        a.P.bc_rho[i] = c->face[i].rho;
        for (int k = 0; k < 3; ++k) a.P.bc_vel[i][k] = c->face[i].vel[k];
        if (c->face[i].type != 0) a.has_bc = 1;
        A.bc_psi_type[i] = c->bc_psi_type[i];
        A.bc_psi_val[i] = c->bc_psi_val[i];
    }
    A.recA = c->d_recA; A.recB = c->d_recB; A.recC = c->d_recC;
    A.rho_r = c->d_rho_r; A.rho_b = c->d_rho_b; A.psi = c->d_psi;
    A.psi_solid = (float)c->psi_solid;
    A.CapA = (float)c->CapA;
    // :100-108, Python float arithmetic, one rounding to f32 where the kernel captures them
    const double wl = 1.0 / (c->niu_l / (1.0 / 3.0) + 0.5);
    const double wg = 1.0 / (c->niu_g / (1.0 / 3.0) + 0.5);
    const double lg0 = 2 * wl * wg / (wl + wg);
    const double l1 = 2 * (wl - lg0) * 10;
    const double l2 = -l1 / 0.2;
    const double g1 = 2 * (lg0 - wg) * 10;
    const double g2 = g1 / 0.2;
    A.wl = (float)wl; A.wg = (float)wg; A.lg0 = (float)lg0;
    A.l1 = (float)l1; A.l2 = (float)l2; A.g1 = (float)g1; A.g2 = (float)g2;
}

int launch_main2(lbm2p_ctx *c, int mode, const Step2Args &A, cudaStream_t st) {
    cudaError_t e = c->cfg.strict ? lbm2p_strict::launch_main(mode, A, c->block, st)
                                  : lbm2p_fast::launch_main(mode, A, c->block, st);
    if (e != cudaSuccess) FAIL2(c, -2, "kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return 0;
}

int launch_colour2(lbm2p_ctx *c, const Step2Args &A, cudaStream_t st) {
    cudaError_t e = c->cfg.strict ? lbm2p_strict::launch_colour(A, c->block, st)
                                  : lbm2p_fast::launch_colour(A, c->block, st);
    if (e != cudaSuccess) FAIL2(c, -2, "kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return 0;
}

int ensure_F2(lbm2p_ctx *c) {
    if (c->d_F != nullptr) return 0;
    CU2(c, cudaMalloc(&c->d_F, c->N * 19 * sizeof(float)));
    k2p_fill_F<<<nblocks(c->N * 19, 256), 256, 0, c->stream>>>(c->d_solid, c->d_F, c->N);
    CU2(c, cudaGetLastError());
    c->launches++;
    return 0;
}

// bring rho_r, rho_b, psi (colour pass) and, if asked, rho, v, F up to the last completed step
int sync2(lbm2p_ctx *c, bool need_macro, bool need_F) {
    if (!c->inited) FAIL2(c, -4, "lbm2p_init has not been called");
    CU2(c, cudaSetDevice(c->cfg.device));
    if (need_F) {
        const bool fresh = c->d_F == nullptr;
        int r = ensure_F2(c);
        if (r) return r;
        if (fresh && c->pipe_valid) c->F_valid = false;
    }
    if (!c->pipe_valid) return 0;
    Step2Args A;
    fill2(c, A, c->d_f[c->cur], nullptr);
    if (!c->colour_valid) {
        int r = launch_colour2(c, A, c->stream);
        if (r) return r;
        c->colour_valid = true;
    }
    if ((need_macro && !c->macro_valid) || (need_F && !c->F_valid)) {
        A.a.F = need_F ? c->d_F : nullptr;
        int r = launch_main2(c, MODE_EXTRACT, A, c->stream);
        if (r) return r;
        c->macro_valid = true;
        if (need_F) c->F_valid = true;
    }
    return 0;
}

int out2(lbm2p_ctx *c, void *dst, const void *src, size_t bytes) {
    CU2(c, cudaStreamSynchronize(c->stream));
    CU2(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return 0;
}

}  // namespace

extern "C" {

const char *lbm2p_last_error(const lbm2p_ctx *ctx) { return ctx ? ctx->err.c_str() : g2_create_error.c_str(); }

int lbm2p_create(const lbm2p_config *cfg, lbm2p_ctx **out) {
    if (!cfg || !out) { g2_create_error = "null argument"; return -1; }
    *out = nullptr;
    if (cfg->nx < 2 || cfg->ny < 2 || cfg->nz < 2) { g2_create_error = "extents must be >= 2"; return -1; }
    const size_t N = (size_t)cfg->nx * cfg->ny * cfg->nz;
    const size_t nzp = ((size_t)cfg->nz + 31) / 32 * 32;
    if ((size_t)cfg->nx * cfg->ny * 19 * nzp + 2 * ((size_t)cfg->ny + 2) * 19 * nzp >= ((size_t)1 << 32)) {
        g2_create_error = "two-phase lattice per context limited by the 32-bit population index";
        return -1;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g2_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                          " (this library has no CPU fallback)";
        return -2;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g2_create_error = "bad device ordinal"; return -1; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g2_create_error = cudaGetErrorString(e); return -2; }
    lbm2p_ctx *c = new lbm2p_ctx();
    c->cfg = *cfg;
    c->N = N;
    if (const char *b = getenv("LBM3D_BLOCK")) {
        int v = atoi(b);
        if (v >= 32 && v <= 256 && v % 32 == 0) c->block = v;
    }
    for (int i = 0; i < 19; ++i)
        for (int j = 0; j < 19; ++j) c->invM[i * 19 + j] = (float)kInvM[i][j];
    e = cudaMalloc(&c->d_solid, N);
    if (e == cudaSuccess) e = cudaMemset(c->d_solid, 0, N);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_psi0, N * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->d_psi0, 0, N * sizeof(float));
    if (e != cudaSuccess) {
        g2_create_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        cudaFree(c->d_solid);
        delete c;
        return -3;
    }
    *out = c;
    return 0;
}

int lbm2p_destroy(lbm2p_ctx *c) {
    CTX2(c);
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    free2(c);
    cudaFree(c->d_solid);
    cudaFree(c->d_psi0);
    delete c;
    return 0;
}

int lbm2p_set_geometry(lbm2p_ctx *c, const int8_t *solid) {
    CTX2(c);
    if (!solid) FAIL2(c, -1, "null geometry");
    if (c->inited) FAIL2(c, -4, "geometry must be set before lbm2p_init");
    CU2(c, cudaSetDevice(c->cfg.device));
    CU2(c, cudaMemcpy(c->d_solid, solid, c->N, cudaMemcpyDefault));
    k_binarize<<<nblocks(c->N, 256), 256>>>(c->d_solid, c->N);
    CU2(c, cudaGetLastError());
    c->launches++;
    return 0;
}

int lbm2p_set_phase(lbm2p_ctx *c, const float *psi) {
    CTX2(c);
    if (!psi) FAIL2(c, -1, "null phase field");
    if (c->inited) FAIL2(c, -4, "the phase field must be set before lbm2p_init (use lbm2p_set_state afterwards)");
    CU2(c, cudaSetDevice(c->cfg.device));
    CU2(c, cudaMemcpy(c->d_psi0, psi, c->N * sizeof(float), cudaMemcpyDefault));
    return 0;
}

int lbm2p_set_fluid(lbm2p_ctx *c, double niu_l, double niu_g, double psi_solid, double CapA) {
    CTX2(c);
    c->niu_l = niu_l; c->niu_g = niu_g; c->psi_solid = psi_solid; c->CapA = CapA;
    return 0;
}

int lbm2p_set_force(lbm2p_ctx *c, const float force[3]) {
    CTX2(c);
    if (!force) FAIL2(c, -1, "null force");
    for (int k = 0; k < 3; ++k) c->force[k] = force[k];
    return 0;
}

int lbm2p_set_bc(lbm2p_ctx *c, int face, int type, float rho, const float vel[3]) {
    CTX2(c);
    if (face < 0 || face > 5 || type < 0 || type > 2) FAIL2(c, -1, "bad face/type");
    if (c->inited) FAIL2(c, -4, "boundary conditions are fixed at lbm2p_init");
    c->face[face].type = type;
    if (type == 1) c->face[face].rho = rho;
    if (type == 2 && vel) for (int k = 0; k < 3; ++k) c->face[face].vel[k] = vel[k];
    return 0;
}

int lbm2p_set_psi_bc(lbm2p_ctx *c, int face, int type, float psi) {
    CTX2(c);
    if (face < 0 || face > 5 || type < 0 || type > 1) FAIL2(c, -1, "bad face/type");
    if (c->inited) FAIL2(c, -4, "boundary conditions are fixed at lbm2p_init");
    c->bc_psi_type[face] = type;
    c->bc_psi_val[face] = psi;
    return 0;
}

int lbm2p_set_inverse_matrix(lbm2p_ctx *c, const float invM[361]) {
    CTX2(c);
    if (!invM) FAIL2(c, -1, "null matrix");
    memcpy(c->invM, invM, sizeof c->invM);
    if (c->inited && c->cfg.strict) {
        CU2(c, cudaSetDevice(c->cfg.device));
        CU2(c, lbm2p_strict::set_inverse_matrix(c->invM));
    }
    return 0;
}

int lbm2p_init(lbm2p_ctx *c) {
    CTX2(c);
    CU2(c, cudaSetDevice(c->cfg.device));
    const int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
    const size_t N = c->N, plane = (size_t)ny * nz;
    free2(c);
    c->inited = false;
    GeoParams g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.halo_x = 0;
    g.xface0 = 0; g.xface1 = nx - 1;
    for (int i = 0; i < 6; ++i) { g.bc_type[i] = c->face[i].type; g.bc_psi_type[i] = c->bc_psi_type[i]; }
    g.two_phase = 1;
    CU2(c, cudaMalloc(&c->d_scalar, 16));
    CU2(c, cudaMalloc(&c->d_flags, N * sizeof(uint32_t)));
    CU2(c, cudaMalloc(&c->d_cls, N));
    k_build_flags<<<nblocks(N, 256), 256>>>(g, c->d_solid, c->d_flags, c->d_cls);
    CU2(c, cudaGetLastError());
    c->launches++;
    {
        auto it = thrust::make_transform_iterator((const int8_t *)c->d_solid, IsFluid());
        uint32_t *d_out = nullptr;
        CU2(c, cudaMalloc(&d_out, sizeof(uint32_t)));
        size_t tmp_bytes = 0;
        void *tmp = nullptr;
        cub::DeviceReduce::Sum(nullptr, tmp_bytes, it, d_out, N);
        CU2(c, cudaMalloc(&tmp, tmp_bytes));
        cudaError_t e = cub::DeviceReduce::Sum(tmp, tmp_bytes, it, d_out, N);
        uint32_t h = 0;
        cudaError_t e2 = cudaMemcpy(&h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        cudaFree(tmp); cudaFree(d_out);
        CU2(c, e); CU2(c, e2);
        c->nf = h;
        c->spec = (double)h >= 0.75 * (double)N ? 1 : 0;
        if (const char *sp = getenv("LBM3D_SPEC")) c->spec = atoi(sp) ? 1 : 0;
    }
    c->nzp = (uint32_t)((nz + 31) / 32 * 32);
    c->prow = 19 * c->nzp;
    c->fsize = (size_t)nx * ny * c->prow;
    c->pad = (((size_t)ny + 1) * c->prow + 2 + 31) / 32 * 32;
    const size_t fbytes = (c->fsize + 2 * c->pad) * sizeof(float);
    for (int b = 0; b < 2; ++b) {
        CU2(c, cudaMalloc(&c->d_fbase[b], fbytes));
        CU2(c, cudaMemset(c->d_fbase[b], 0, fbytes));
        c->d_f[b] = c->d_fbase[b] + c->pad;
    }
    // node-linear arrays with a guard band of a plane + a row
    c->npad = (plane + nz + 2 + 31) / 32 * 32;
    const size_t nlin = N + 2 * c->npad;
    CU2(c, cudaMalloc(&c->d_recA, N * sizeof(float4)));
    CU2(c, cudaMalloc(&c->d_recB, N * sizeof(float2)));
    CU2(c, cudaMalloc(&c->d_recC, N * sizeof(float4)));
    CU2(c, cudaMemset(c->d_recA, 0, N * sizeof(float4)));
    CU2(c, cudaMemset(c->d_recB, 0, N * sizeof(float2)));
    CU2(c, cudaMemset(c->d_recC, 0, N * sizeof(float4)));
    CU2(c, cudaMalloc(&c->d_psibase, nlin * sizeof(float)));
    CU2(c, cudaMemset(c->d_psibase, 0, nlin * sizeof(float)));
    c->d_psi = c->d_psibase + c->npad;
    CU2(c, cudaMalloc(&c->d_rho_r, N * sizeof(float)));
    CU2(c, cudaMalloc(&c->d_rho_b, N * sizeof(float)));
    CU2(c, cudaMalloc(&c->d_rho, N * sizeof(float)));
    CU2(c, cudaMalloc(&c->d_v, N * 3 * sizeof(float)));
    k2p_init_state<<<nblocks(N, 256), 256>>>(c->d_solid, c->d_psi0, N, (float)c->psi_solid, c->d_psi, c->d_rho_r,
                                              c->d_rho_b, c->d_rho, c->d_v);
    CU2(c, cudaGetLastError());
    c->launches++;
    const size_t fs[6] = {plane, plane, (size_t)nx * nz, (size_t)nx * nz, (size_t)nx * ny, (size_t)nx * ny};
    size_t tot = 0;
    for (int i = 0; i < 6; ++i) { c->vbc_off[i] = (uint32_t)tot; tot += fs[i]; }
    CU2(c, cudaMalloc(&c->d_vbc, tot * 3 * sizeof(float)));
    CU2(c, cudaMemset(c->d_vbc, 0, tot * 3 * sizeof(float)));
    if (c->cfg.strict) CU2(c, lbm2p_strict::set_inverse_matrix(c->invM));
    CU2(c, cudaDeviceSynchronize());
    c->cur = 0;
    c->pipe_valid = false;
    c->colour_valid = true;
    c->macro_valid = true;
    c->F_valid = true;
    c->inited = true;
    return 0;
}

int lbm2p_step(lbm2p_ctx *c, int nsteps, void *cuda_stream) {
    CTX2(c);
    if (!c->inited) FAIL2(c, -4, "lbm2p_init has not been called");
    if (nsteps < 0) FAIL2(c, -1, "nsteps < 0");
    if (nsteps == 0) return 0;
    CU2(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    c->stream = st;
    Step2Args A;
    if (!c->pipe_valid) {
        // collision of the first step from the user-visible state (:302-372)
        fill2(c, A, nullptr, c->d_f[c->cur]);
        A.a.F = c->d_F;
        int r = launch_main2(c, MODE_COLLIDE, A, st);
        if (r) return r;
        c->pipe_valid = true;
        c->colour_valid = false;
        nsteps -= 1;
    }
    for (int it = 0; it < nsteps; ++it) {
        fill2(c, A, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
        if (!c->colour_valid) {
            int r = launch_colour2(c, A, st);       // rho_r, rho_b, psi of the step just collided
            if (r) return r;
        }
        int r = launch_main2(c, MODE_STEP, A, st);  // its stream/BC/macro + the next collision
        if (r) return r;
        c->cur ^= 1;
        c->colour_valid = false;
    }
    c->macro_valid = false;
    c->F_valid = false;
    return 0;
}

int64_t lbm2p_launch_count(const lbm2p_ctx *c) { return c ? c->launches : -1; }

int lbm2p_synchronize(lbm2p_ctx *c) {
    CTX2(c);
    CU2(c, cudaSetDevice(c->cfg.device));
    CU2(c, cudaStreamSynchronize(c->stream));
    return 0;
}

#define GETTER(name, need_macro, need_F, src, count)                                           \
    int name(lbm2p_ctx *c, float *dst) {                                                       \
        CTX2(c);                                                                               \
        if (!dst) FAIL2(c, -1, "null destination");                                            \
        int r = sync2(c, need_macro, need_F);                                                  \
        if (r) return r;                                                                       \
        return out2(c, dst, src, (count) * sizeof(float));                                     \
    }
GETTER(lbm2p_get_rho, true, false, c->d_rho, c->N)
GETTER(lbm2p_get_v, true, false, c->d_v, c->N * 3)
GETTER(lbm2p_get_F, true, true, c->d_F, c->N * 19)
GETTER(lbm2p_get_rho_r, false, false, c->d_rho_r, c->N)
GETTER(lbm2p_get_rho_b, false, false, c->d_rho_b, c->N)
#undef GETTER

int lbm2p_get_psi(lbm2p_ctx *c, float *dst) {
    CTX2(c);
    if (!dst) FAIL2(c, -1, "null destination");
    int r = sync2(c, false, false);
    if (r) return r;
    float *tmp = nullptr;
    CU2(c, cudaMalloc(&tmp, c->N * sizeof(float)));
    k2p_merge_psi<<<nblocks(c->N, 256), 256, 0, c->stream>>>(c->d_solid, c->d_psi0, c->d_psi, tmp, c->N);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(dst, tmp, c->N * sizeof(float), cudaMemcpyDefault);
    cudaFree(tmp);
    CU2(c, e);
    c->launches++;
    return 0;
}

int lbm2p_get_solid(lbm2p_ctx *c, int8_t *dst) {
    CTX2(c);
    if (!dst) FAIL2(c, -1, "null destination");
    CU2(c, cudaSetDevice(c->cfg.device));
    CU2(c, cudaMemcpy(dst, c->d_solid, c->N, cudaMemcpyDefault));
    return 0;
}

int lbm2p_set_state(lbm2p_ctx *c, const float *F, const float *rho, const float *v, const float *psi,
                    const float *rho_r, const float *rho_b) {
    CTX2(c);
    if (!F || !rho || !v || !psi || !rho_r || !rho_b) FAIL2(c, -1, "null field");
    if (!c->inited) FAIL2(c, -4, "lbm2p_init has not been called");
    CU2(c, cudaSetDevice(c->cfg.device));
    CU2(c, cudaStreamSynchronize(c->stream));
    int r = ensure_F2(c);
    if (r) return r;
    CU2(c, cudaStreamSynchronize(c->stream));
    CU2(c, cudaMemcpy(c->d_F, F, c->N * 19 * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_rho, rho, c->N * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_v, v, c->N * 3 * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_psi0, psi, c->N * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_psi, psi, c->N * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_rho_r, rho_r, c->N * sizeof(float), cudaMemcpyDefault));
    CU2(c, cudaMemcpy(c->d_rho_b, rho_b, c->N * sizeof(float), cudaMemcpyDefault));
    k2p_bake_psi<<<nblocks(c->N, 256), 256>>>(c->d_solid, c->d_psi, (float)c->psi_solid, c->N);
    CU2(c, cudaGetLastError());
    CU2(c, cudaDeviceSynchronize());
    c->launches++;
    c->pipe_valid = false;
    c->colour_valid = true;
    c->macro_valid = true;
    c->F_valid = true;
    return 0;
}

int lbm2p_get_max_v(lbm2p_ctx *c, float *out) {
    CTX2(c);
    if (!out) FAIL2(c, -1, "null destination");
    int r = sync2(c, true, false);
    if (r) return r;
    const int init_bits = 0x80000000;
    CU2(c, cudaMemcpyAsync(c->d_scalar, &init_bits, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_max_v<<<148 * 8, 256, 0, c->stream>>>(c->d_v, c->N, c->d_scalar);
    CU2(c, cudaGetLastError());
    c->launches++;
    int bits = 0;
    CU2(c, cudaMemcpyAsync(&bits, c->d_scalar, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU2(c, cudaStreamSynchronize(c->stream));
    float v;
    memcpy(&v, &bits, sizeof v);
    *out = bits < 0 ? -1e10f : v;
    return 0;
}

}  // extern "C"
