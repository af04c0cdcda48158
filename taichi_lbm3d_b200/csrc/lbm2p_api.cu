// C-ABI layer of the two-phase colour-gradient solver (include/lbm3d_2phase.h).
// Reference: 2phase/lbm_solver_3d_2phase.py (line numbers below).  Dense storage; one GPU or an
// x-slab of a multi-GPU run (ghost planes + two halo exchanges per step, see lbm2p_run_slab).
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
#include "lbm_nccl.cuh"
#include "lbm_sparse_build.cuh"

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
    int gather = 0;               // dense colour pass gathers per node (few BULK nodes)
    // device
    int8_t *d_solid = nullptr;
    float *d_psi0 = nullptr;          // input phase field (kept for the solid nodes of get_psi)
    uint32_t *d_flags = nullptr;
    uint8_t *d_cls = nullptr;
    float *d_fbase[2] = {nullptr, nullptr}, *d_f[2] = {nullptr, nullptr};
    float4 *d_uq = nullptr, *d_recC = nullptr;     // colour record: (v, +-q) and, at the interface, (C, 1/|C|)
    float2 *d_rrb[2] = {nullptr, nullptr};         // (rho_r, rho_b): colour pass reads [rcur], writes [rcur ^ 1]
    int rcur = 0;
    float *d_psibase = nullptr, *d_psi = nullptr;
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
    // sparse storage (cfg.reserved & LBM2P_SPARSE): compact fluid list + pull table; the colour
    // records, rho_r, rho_b, psi and the populations are indexed by stored node
    bool sparse = false;
    SparseTables sp;
    // x-slab (cfg.reserved bit0): planes 0 and nx-1 are ghost planes
    bool halo = false;
    int xface0 = 0, xface1 = 0;            // local x of the global x faces (-1: not in this slab)
    uint32_t row_first = 0, row_count = 0;  // z-rows updated by a step
    void *comm = nullptr;                  // ncclComm_t
    int comm_world = 1, comm_rank = 0;
    bool comm_ready = false;
    float *d_send[2] = {nullptr, nullptr}, *d_recv[2] = {nullptr, nullptr};
    cudaStream_t comm_stream = nullptr;    // highest priority: the exchanges run beside the interior kernels
    cudaEvent_t ev_cb = nullptr, ev_xb = nullptr, ev_mb = nullptr, ev_xa = nullptr;
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
                               float psi_solid, float *psi, float2 *rrb, float *rho, float *v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (solid[i] == 0) {
        const float p = psi0[i];
        psi[i] = p;
        const float rr = (p + 1.0f) / 2.0f;
        rrb[i] = make_float2(rr, 1.0f - rr);
        rho[i] = 1.0f;
    } else {
        psi[i] = psi_solid;
        rrb[i] = make_float2(0.f, 0.f);
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

// sparse storage: init :173-186 on the compact arrays; psi0 is the dense input phase field
__global__ void k2p_init_compact(const uint32_t *__restrict__ lin, const float *__restrict__ psi0, size_t nf,
                                 float *psi, float2 *rrb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const float p = psi0[lin[i]];
    psi[i] = p;
    const float rr = (p + 1.0f) / 2.0f;
    rrb[i] = make_float2(rr, 1.0f - rr);
}
// one colour of the interleaved (rho_r, rho_b) array as its own array (getters), and back (set_state)
__global__ void k2p_split(const float2 *__restrict__ rrb, int which, size_t n, float *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = which ? rrb[i].y : rrb[i].x;
}
__global__ void k2p_join(const float *__restrict__ rr, const float *__restrict__ rb, const uint32_t *__restrict__ lin,
                         size_t n, float2 *rrb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t j = lin ? lin[i] : i;
    rrb[i] = make_float2(rr[j], rb[j]);
}
__global__ void k2p_init_macro(const int8_t *__restrict__ solid, size_t n, float *rho, float *v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rho[i] = solid[i] == 0 ? 1.0f : 0.f;
    v[3 * i] = 0.f; v[3 * i + 1] = 0.f; v[3 * i + 2] = 0.f;
}
// compact -> dense (getters): `out` is pre-filled with what solid nodes show
__global__ void k2p_scatter(const uint32_t *__restrict__ lin, const float *__restrict__ src, size_t nf, float *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) out[lin[i]] = src[i];
}
// dense -> compact (set_state)
__global__ void k2p_gather(const uint32_t *__restrict__ lin, const float *__restrict__ src, size_t nf, float *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) out[i] = src[lin[i]];
}

void free2(lbm2p_ctx *c) {
    free_sparse_tables(c->sp);
    cudaFree(c->d_flags); cudaFree(c->d_cls); cudaFree(c->d_fbase[0]); cudaFree(c->d_fbase[1]);
    cudaFree(c->d_uq); cudaFree(c->d_recC); cudaFree(c->d_psibase); cudaFree(c->d_rrb[0]); cudaFree(c->d_rrb[1]);
    cudaFree(c->d_rho); cudaFree(c->d_v); cudaFree(c->d_F); cudaFree(c->d_vbc); cudaFree(c->d_scalar);
    c->d_flags = nullptr; c->d_cls = nullptr; c->d_fbase[0] = c->d_fbase[1] = nullptr;
    c->d_uq = c->d_recC = nullptr; c->d_psibase = nullptr; c->d_rrb[0] = c->d_rrb[1] = nullptr;
    c->d_rho = c->d_v = c->d_F = nullptr; c->d_vbc = nullptr; c->d_scalar = nullptr;
}

void fill2(const lbm2p_ctx *c, Step2Args &A, const float *fin, float *fout) {
    memset(&A, 0, sizeof A);
    StepArgs &a = A.a;
    static const int e[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
        {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1},
        {1, 0, -1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};
    const long long sy = (long long)c->prow, sx = (long long)c->cfg.ny * sy;
    const size_t pstride = c->sparse ? c->sp.stride : (size_t)c->nzp;
    for (int s = 0; s < 19; ++s) {
        a.pown[s] = fin ? fin + (size_t)s * pstride : nullptr;
        a.ppull[s] = fin ? a.pown[s] - (e[s][0] * sx + e[s][1] * sy + e[s][2]) : nullptr;
        a.pout[s] = fout ? fout + (size_t)s * pstride : nullptr;
    }
    a.stride = 0;
    A.sparse = c->sparse ? 1 : 0;
    A.gather = c->gather;
    if (c->sparse) {
        a.stride = c->sp.stride;
        a.first = c->sp.own_first;
        a.count = c->sp.own_count;
        a.lin = c->sp.d_lin;
        a.compressed = 1;
        for (int k = 0; k < 8; ++k) a.rb8[k] = c->sp.d_rb8 + (size_t)k * c->sp.stride;
        a.blk = c->sp.d_blk;
        for (int k = 0; k < 18; ++k) a.exc[k] = c->sp.d_exc + (size_t)k * c->sp.exc_stride;
    }
    a.row_first = c->row_first;
    a.row_count = c->row_count;
    a.row_split = 0xFFFFFFFFu; a.row_skip = 0;
    a.nb1 = 0xFFFFFFFFu; a.first2 = 0; a.count2 = 0;
    a.halo_x = c->halo ? 1 : 0;
    a.prow = c->prow;
    a.spec = c->spec;
    a.nx = c->cfg.nx; a.ny = c->cfg.ny; a.nz = c->cfg.nz;
    a.flags = c->sparse ? c->sp.d_flags : c->d_flags;
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
    A.uq = c->d_uq; A.recC = c->d_recC;
    A.rrb = c->d_rrb[c->rcur]; A.rrb_out = c->d_rrb[c->rcur ^ 1];
    A.psi = c->d_psi;
    if (!c->sparse) {
        const long long py = c->cfg.nz, px = (long long)c->cfg.ny * py;
        for (int s = 0; s < 19; ++s) A.psi_nb[s] = c->d_psi + (e[s][0] * px + e[s][1] * py + e[s][2]);
    }
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
    cudaError_t e;
    if (c->sparse)
        e = c->cfg.strict ? lbm2p_strict::launch_main_sparse(mode, A, st) : lbm2p_fast::launch_main_sparse(mode, A, st);
    else
        e = c->cfg.strict ? lbm2p_strict::launch_main(mode, A, c->block, st)
                          : lbm2p_fast::launch_main(mode, A, c->block, st);
    if (e != cudaSuccess) FAIL2(c, -2, "kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return 0;
}

int launch_colour2(lbm2p_ctx *c, const Step2Args &A, cudaStream_t st) {
    cudaError_t e;
    if (c->sparse)
        e = c->cfg.strict ? lbm2p_strict::launch_colour_sparse(A, st) : lbm2p_fast::launch_colour_sparse(A, st);
    else
        e = c->cfg.strict ? lbm2p_strict::launch_colour(A, c->block, st)
                          : lbm2p_fast::launch_colour(A, c->block, st);
    if (e != cudaSuccess) FAIL2(c, -2, "kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return 0;
}

// the colour pass of a step is complete (it may take several launches over plane ranges): what it
// wrote is the current (rho_r, rho_b) from here on
void colour_done(lbm2p_ctx *c) {
    c->rcur ^= 1;
    c->colour_valid = true;
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
        colour_done(c);
        fill2(c, A, c->d_f[c->cur], nullptr);
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
    if (N >= ((size_t)1 << 32)) { g2_create_error = "lattice per context limited to 2^32 nodes"; return -1; }
    if (!(cfg->reserved & LBM2P_SPARSE) &&
        (size_t)cfg->nx * cfg->ny * 19 * nzp + 2 * ((size_t)cfg->ny + 2) * 19 * nzp >= ((size_t)1 << 32)) {
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
    c->halo = (cfg->reserved & LBM2P_HALO_X) != 0;
    c->sparse = (cfg->reserved & LBM2P_SPARSE) != 0;
    if (c->halo && cfg->nx < 3) { g2_create_error = "an x-slab needs nx >= 3 (two ghost planes)"; delete c; return -1; }
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
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_send[i]); cudaFree(c->d_recv[i]); }
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    for (cudaEvent_t ev : {c->ev_cb, c->ev_xb, c->ev_mb, c->ev_xa})
        if (ev) cudaEventDestroy(ev);
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
    g.halo_x = c->halo ? 1 : 0;
    if (c->halo) {
        // the global x faces are the first / last OWNED planes of the slabs that hold them
        c->xface0 = (c->cfg.reserved & LBM2P_HOLDS_X0) ? 1 : -1;
        c->xface1 = (c->cfg.reserved & LBM2P_HOLDS_X1) ? nx - 2 : -1;
        c->row_first = (uint32_t)ny;
        c->row_count = (uint32_t)(ny * (nx - 2));
    } else {
        c->xface0 = 0; c->xface1 = nx - 1;
        c->row_first = 0;
        c->row_count = (uint32_t)(nx * ny);
    }
    g.xface0 = c->xface0; g.xface1 = c->xface1;
    for (int i = 0; i < 6; ++i) { g.bc_type[i] = c->face[i].type; g.bc_psi_type[i] = c->bc_psi_type[i]; }
    g.two_phase = 1;
    g.vel_in_place = 0;
    CU2(c, cudaMalloc(&c->d_scalar, 16));
    CU2(c, cudaMalloc(&c->d_rho, N * sizeof(float)));
    CU2(c, cudaMalloc(&c->d_v, N * 3 * sizeof(float)));
    if (c->sparse) {
        // compact fluid list + pull table (lbm_sparse_build.cuh); link words carry the face and
        // wetting bits of the two-phase kernels
        std::string msg;
        cudaError_t e = build_sparse_tables(g, c->d_solid, true, c->sp, msg);
        c->launches += c->sp.launches;
        if (!msg.empty()) FAIL2(c, -1, "%s", msg.c_str());
        CU2(c, e);
        c->nf = c->sp.nf;
        const size_t st = c->sp.stride;
        c->fsize = 19 * st;
        c->pad = 0;
        for (int b = 0; b < 2; ++b) {
            CU2(c, cudaMalloc(&c->d_fbase[b], c->fsize * sizeof(float)));
            CU2(c, cudaMemset(c->d_fbase[b], 0, c->fsize * sizeof(float)));
            c->d_f[b] = c->d_fbase[b];
        }
        CU2(c, cudaMalloc(&c->d_uq, st * sizeof(float4)));
        CU2(c, cudaMalloc(&c->d_recC, st * sizeof(float4)));
        CU2(c, cudaMemset(c->d_uq, 0, st * sizeof(float4)));
        CU2(c, cudaMemset(c->d_recC, 0, st * sizeof(float4)));
        CU2(c, cudaMalloc(&c->d_psibase, st * sizeof(float)));
        CU2(c, cudaMemset(c->d_psibase, 0, st * sizeof(float)));
        c->d_psi = c->d_psibase;
        for (int b = 0; b < 2; ++b) {
            CU2(c, cudaMalloc(&c->d_rrb[b], st * sizeof(float2)));
            CU2(c, cudaMemset(c->d_rrb[b], 0, st * sizeof(float2)));
        }
        if (c->nf) {
            k2p_init_compact<<<nblocks(c->nf, 256), 256>>>(c->sp.d_lin, c->d_psi0, c->nf, c->d_psi, c->d_rrb[0]);
            CU2(c, cudaGetLastError());
        }
        k2p_init_macro<<<nblocks(N, 256), 256>>>(c->d_solid, N, c->d_rho, c->d_v);
        CU2(c, cudaGetLastError());
        c->launches += 2;
    } else {
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
        // the colour pass hands terms between lanes when most fluid nodes are BULK (no solid link,
        // no face); in a porous medium nearly every node gathers for itself anyway
        auto itb = thrust::make_transform_iterator((const uint8_t *)c->d_cls, IsBulk());
        CU2(c, cudaMalloc(&d_out, sizeof(uint32_t)));
        tmp = nullptr; tmp_bytes = 0;
        cub::DeviceReduce::Sum(nullptr, tmp_bytes, itb, d_out, N);
        CU2(c, cudaMalloc(&tmp, tmp_bytes));
        e = cub::DeviceReduce::Sum(tmp, tmp_bytes, itb, d_out, N);
        uint32_t hb = 0;
        e2 = cudaMemcpy(&hb, d_out, sizeof hb, cudaMemcpyDeviceToHost);
        cudaFree(tmp); cudaFree(d_out);
        CU2(c, e); CU2(c, e2);
        c->gather = (double)hb < 0.5 * (double)h ? 1 : 0;
        if (const char *g = getenv("LBM3D_COLOUR_GATHER")) c->gather = atoi(g) ? 1 : 0;
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
    CU2(c, cudaMalloc(&c->d_uq, N * sizeof(float4)));
    CU2(c, cudaMalloc(&c->d_recC, N * sizeof(float4)));
    CU2(c, cudaMemset(c->d_uq, 0, N * sizeof(float4)));
    CU2(c, cudaMemset(c->d_recC, 0, N * sizeof(float4)));
    CU2(c, cudaMalloc(&c->d_psibase, nlin * sizeof(float)));
    CU2(c, cudaMemset(c->d_psibase, 0, nlin * sizeof(float)));
    c->d_psi = c->d_psibase + c->npad;
    for (int b = 0; b < 2; ++b) {
        CU2(c, cudaMalloc(&c->d_rrb[b], N * sizeof(float2)));
        CU2(c, cudaMemset(c->d_rrb[b], 0, N * sizeof(float2)));
    }
    k2p_init_state<<<nblocks(N, 256), 256>>>(c->d_solid, c->d_psi0, N, (float)c->psi_solid, c->d_psi, c->d_rrb[0],
                                              c->d_rho, c->d_v);
    CU2(c, cudaGetLastError());
    c->launches++;
    }
    const size_t fs[6] = {plane, plane, (size_t)nx * nz, (size_t)nx * nz, (size_t)nx * ny, (size_t)nx * ny};
    size_t tot = 0;
    for (int i = 0; i < 6; ++i) { c->vbc_off[i] = (uint32_t)tot; tot += fs[i]; }
    CU2(c, cudaMalloc(&c->d_vbc, tot * 3 * sizeof(float)));
    CU2(c, cudaMemset(c->d_vbc, 0, tot * 3 * sizeof(float)));
    if (c->cfg.strict) CU2(c, lbm2p_strict::set_inverse_matrix(c->invM));
    CU2(c, cudaDeviceSynchronize());
    c->cur = 0;
    c->rcur = 0;
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
    if (c->halo) FAIL2(c, -4, "an x-slab context is stepped with lbm2p_run_slab (halo exchange inside)");
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
            colour_done(c);
            fill2(c, A, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
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

// ---- multi-GPU x-slabs -----------------------------------------------------------------------
// A slab context (cfg.reserved & LBM2P_HALO_X) owns planes 1..nx-2; planes 0 and nx-1 mirror
// the neighbours' boundary planes.  Two exchanges per step (include/lbm3d_2phase.h):
//   stage 0, after the main pass:   5 face-crossing populations of f* + the colour records
//   stage 1, after the colour pass: psi
namespace {

// nodes of halo plane 0 (left ghost), 1 (first owned), 2 (last owned), 3 (right ghost) and where
// the plane starts in the node-indexed arrays (dense: node-linear; sparse: compact list)
size_t plane_nodes(const lbm2p_ctx *c, int plane) {
    return c->sparse ? (size_t)c->sp.plane_count[plane] : (size_t)c->cfg.ny * c->cfg.nz;
}
size_t plane_start(const lbm2p_ctx *c, int plane) {
    if (c->sparse) return c->sp.plane_first[plane];
    const int x = plane == 0 ? 0 : (plane == 1 ? 1 : (plane == 2 ? c->cfg.nx - 2 : c->cfg.nx - 1));
    return (size_t)x * c->cfg.ny * c->cfg.nz;
}
size_t stage_floats(int stage, size_t nodes) {
    return (stage == 0 ? 13 : 3) * nodes;      // f* (5) + uq (4) + recC (4)  |  psi (1) + rho_r, rho_b (2)
}
size_t halo_floats(const lbm2p_ctx *c, int stage) {
    size_t most = 0;
    for (int pl = 0; pl < 4; ++pl) most = plane_nodes(c, pl) > most ? plane_nodes(c, pl) : most;
    return stage_floats(stage, most);
}

int pack2(lbm2p_ctx *c, int stage, int side, float *dst, cudaStream_t st) {
    const int plane = side == 0 ? 1 : 2;                       // boundary plane that is sent
    const size_t P = plane_nodes(c, plane), off = plane_start(c, plane);
    if (P == 0) return 0;
    if (stage == 1) {       // what the colour pass just wrote: psi and the current (rho_r, rho_b)
        CU2(c, cudaMemcpyAsync(dst, c->d_psi + off, P * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CU2(c, cudaMemcpyAsync(dst + P, c->d_rrb[c->rcur] + off, P * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    static const HaloDirs kR = {{1, 7, 9, 11, 13}}, kL = {{2, 8, 10, 12, 14}};
    Step2Args A;
    fill2(c, A, c->d_f[c->cur], nullptr);
    if (c->sparse) {
        A.a.nz = 0;
        k_halo_pack<<<nblocks(P, 256), 256, 0, st>>>(A.a, 0u, (uint32_t)off, (uint32_t)P, side == 0 ? kL : kR, dst);
    } else {
        const int x = side == 0 ? 1 : c->cfg.nx - 2;
        k_halo_pack<<<nblocks(P, 256), 256, 0, st>>>(A.a, (uint32_t)(x * c->cfg.ny), 0u, (uint32_t)P, side == 0 ? kL : kR, dst);
    }
    CU2(c, cudaGetLastError());
    c->launches++;
    CU2(c, cudaMemcpyAsync(dst + 5 * P, c->d_uq + off, P * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CU2(c, cudaMemcpyAsync(dst + 9 * P, c->d_recC + off, P * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int unpack2(lbm2p_ctx *c, int stage, int side, const float *src, cudaStream_t st) {
    const int plane = side == 0 ? 0 : 3;                       // ghost plane that is filled
    const size_t P = plane_nodes(c, plane), off = plane_start(c, plane);
    if (P == 0) return 0;
    if (stage == 1) {
        CU2(c, cudaMemcpyAsync(c->d_psi + off, src, P * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CU2(c, cudaMemcpyAsync(c->d_rrb[c->rcur] + off, src + P, P * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    static const HaloDirs kR = {{1, 7, 9, 11, 13}}, kL = {{2, 8, 10, 12, 14}};
    Step2Args A;
    fill2(c, A, nullptr, c->d_f[c->cur]);
    // the left ghost receives what the left neighbour sent to its right (e_x = +1) and v.v.
    if (c->sparse) {
        A.a.nz = 0;
        k_halo_unpack<<<nblocks(P, 256), 256, 0, st>>>(A.a, 0u, (uint32_t)off, (uint32_t)P, side == 0 ? kR : kL, src);
    } else {
        const int x = side == 0 ? 0 : c->cfg.nx - 1;
        k_halo_unpack<<<nblocks(P, 256), 256, 0, st>>>(A.a, (uint32_t)(x * c->cfg.ny), 0u, (uint32_t)P, side == 0 ? kR : kL, src);
    }
    CU2(c, cudaGetLastError());
    c->launches++;
    CU2(c, cudaMemcpyAsync(c->d_uq + off, src + 5 * P, P * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CU2(c, cudaMemcpyAsync(c->d_recC + off, src + 9 * P, P * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    return 0;
}

#define NC2(ctx, call)                                                                         \
    do {                                                                                       \
        int _r = (call);                                                                       \
        if (_r != 0) FAIL2(ctx, -2, "%s failed: %s", #call, g_nccl.GetErrorString(_r));        \
    } while (0)

int exchange2(lbm2p_ctx *c, int stage, cudaStream_t st) {
    int r = pack2(c, stage, 0, c->d_send[0], st);
    if (r) return r;
    r = pack2(c, stage, 1, c->d_send[1], st);
    if (r) return r;
    const float *from_left = c->d_send[1], *from_right = c->d_send[0];     // ring of one slab
    if (c->comm_world > 1) {
        const int left = (c->comm_rank + c->comm_world - 1) % c->comm_world;
        const int right = (c->comm_rank + 1) % c->comm_world;
        NC2(c, g_nccl.GroupStart());
        // posting order matters when left == right (two ranks): pairs match in order
        NC2(c, g_nccl.Send(c->d_send[1], stage_floats(stage, plane_nodes(c, 2)), 7, right, c->comm, st));
        NC2(c, g_nccl.Send(c->d_send[0], stage_floats(stage, plane_nodes(c, 1)), 7, left, c->comm, st));
        NC2(c, g_nccl.Recv(c->d_recv[0], stage_floats(stage, plane_nodes(c, 0)), 7, left, c->comm, st));
        NC2(c, g_nccl.Recv(c->d_recv[1], stage_floats(stage, plane_nodes(c, 3)), 7, right, c->comm, st));
        NC2(c, g_nccl.GroupEnd());
        from_left = c->d_recv[0];
        from_right = c->d_recv[1];
    }
    r = unpack2(c, stage, 0, from_left, st);
    if (r) return r;
    return unpack2(c, stage, 1, from_right, st);
}

// one stage of the slab step: 0 first collision (if the pipeline is empty), 1 colour, 2 main
int stage2(lbm2p_ctx *c, int stage, cudaStream_t st) {
    Step2Args A;
    if (stage == 0) {
        if (c->pipe_valid) return 1;
        fill2(c, A, nullptr, c->d_f[c->cur]);
        A.a.F = c->d_F;
        int r = launch_main2(c, MODE_COLLIDE, A, st);
        if (r) return r;
        c->pipe_valid = true;
        c->colour_valid = false;
    } else if (stage == 1) {
        if (!c->pipe_valid) FAIL2(c, -4, "pipeline not started");
        if (c->colour_valid) return 1;
        fill2(c, A, c->d_f[c->cur], nullptr);
        int r = launch_colour2(c, A, st);
        if (r) return r;
        colour_done(c);
    } else {
        if (!c->pipe_valid || !c->colour_valid) FAIL2(c, -4, "the colour pass of this step has not run");
        fill2(c, A, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
        int r = launch_main2(c, MODE_STEP, A, st);
        if (r) return r;
        c->cur ^= 1;
        c->colour_valid = false;
    }
    c->macro_valid = false;
    c->F_valid = false;
    return 0;
}

}  // namespace

int64_t lbm2p_halo_floats(lbm2p_ctx *c, int stage) {
    if (!c || !c->inited || !c->halo || stage < 0 || stage > 1) return -1;
    return (int64_t)halo_floats(c, stage);
}

int64_t lbm2p_halo_count(lbm2p_ctx *c, int plane) {
    if (!c || !c->inited || !c->halo || plane < 0 || plane > 3) return -1;
    return (int64_t)plane_nodes(c, plane);
}

int lbm2p_halo_pack(lbm2p_ctx *c, int stage, int side, float *dst, void *cuda_stream) {
    CTX2(c);
    if (!c->inited || !c->halo) FAIL2(c, -4, "needs an initialised x-slab context");
    if (!c->pipe_valid) FAIL2(c, -4, "no post-collision state yet (run stage 0 first)");
    if (stage < 0 || stage > 1 || side < 0 || side > 1 || !dst) FAIL2(c, -1, "bad stage/side/destination");
    CU2(c, cudaSetDevice(c->cfg.device));
    return pack2(c, stage, side, dst, (cudaStream_t)cuda_stream);
}

int lbm2p_halo_unpack(lbm2p_ctx *c, int stage, int side, const float *src, void *cuda_stream) {
    CTX2(c);
    if (!c->inited || !c->halo) FAIL2(c, -4, "needs an initialised x-slab context");
    if (stage < 0 || stage > 1 || side < 0 || side > 1 || !src) FAIL2(c, -1, "bad stage/side/source");
    CU2(c, cudaSetDevice(c->cfg.device));
    return unpack2(c, stage, side, src, (cudaStream_t)cuda_stream);
}

int lbm2p_slab_stage(lbm2p_ctx *c, int stage, void *cuda_stream) {
    CTX2(c);
    if (!c->inited || !c->halo) FAIL2(c, -4, "needs an initialised x-slab context");
    if (stage < 0 || stage > 2) FAIL2(c, -1, "stage must be 0 (first collision), 1 (colour) or 2 (main)");
    CU2(c, cudaSetDevice(c->cfg.device));
    c->stream = (cudaStream_t)cuda_stream;
    return stage2(c, stage, c->stream);
}

int lbm2p_comm_init(lbm2p_ctx *c, const void *id128, int world, int rank) {
    CTX2(c);
    if (!c->inited || !c->halo) FAIL2(c, -4, "needs an initialised x-slab context");
    if (world < 1 || rank < 0 || rank >= world) FAIL2(c, -1, "bad world/rank");
    CU2(c, cudaSetDevice(c->cfg.device));
    if (world > 1) {
        std::string err;
        if (!load_nccl(err)) FAIL2(c, -2, "%s", err.c_str());
        NcclId id;
        memset(&id, 0, sizeof id);
        if (id128) memcpy(&id, id128, sizeof id);
        else if (!have_shared_comm(world, rank)) FAIL2(c, -1, "null NCCL id and no communicator yet");
        NC2(c, shared_comm(world, rank, id, &c->comm));
    }
    c->comm_world = world;
    c->comm_rank = rank;
    for (int i = 0; i < 2; ++i) {
        if (!c->d_send[i]) CU2(c, cudaMalloc(&c->d_send[i], (halo_floats(c, 0) + 4) * sizeof(float)));
        if (!c->d_recv[i]) CU2(c, cudaMalloc(&c->d_recv[i], (halo_floats(c, 0) + 4) * sizeof(float)));
    }
    if (!c->comm_stream) {
        int lo = 0, hi = 0;
        CU2(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU2(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
        for (cudaEvent_t *ev : {&c->ev_cb, &c->ev_xb, &c->ev_mb, &c->ev_xa})
            CU2(c, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    }
    c->comm_ready = true;
    return 0;
}

int lbm2p_comm_unique_id(void *out128) {
    std::string err;
    if (!out128) return -1;
    if (!load_nccl(err)) { g2_create_error = err; return -2; }
    NcclId id;
    if (g_nccl.GetUniqueId(&id) != 0) { g2_create_error = "ncclGetUniqueId failed"; return -2; }
    memcpy(out128, &id, sizeof id);
    return 0;
}

// colour pass (kind 0) or main pass (kind 1) over the planes [x0, x1) of the slab
static int launch_planes2(lbm2p_ctx *c, int kind, int x0, int x1, cudaStream_t st) {
    if (x1 <= x0) return 0;
    Step2Args A;
    fill2(c, A, c->d_f[c->cur], kind ? c->d_f[c->cur ^ 1] : nullptr);
    if (c->sparse) {
        A.a.first = c->sp.plane_rank[x0];
        A.a.count = c->sp.plane_rank[x1] - A.a.first;
        if (A.a.count == 0) return 0;
    } else {
        A.a.row_first = (uint32_t)(x0 * c->cfg.ny);
        A.a.row_count = (uint32_t)((x1 - x0) * c->cfg.ny);
    }
    return kind ? launch_main2(c, MODE_STEP, A, st) : launch_colour2(c, A, st);
}

// nsteps iterations of the main loop :626-632 on one x-slab, halo exchanges inside.
// Default: boundary planes first, their exchange on a highest-priority side stream while the interior
// planes are updated (bit-identical to the plain schedule; 2 B200, 2 x 128 x 256^2: 0.590 ms per step
// against 0.704):
//   main stream  colour(interior) . wait XA' . colour(boundary) . main(interior) . wait XB . main(boundary) ...
//   side stream                                  XB = exchange(psi, rho_r, rho_b)   XA = exchange(f*, records)
// (XA' = the exchange of the previous step; the colour pass of interior planes only reads records
// of owned planes, so it runs beside it).  LBM3D_2P_OVERLAP=0 selects the plain schedule: colour ;
// exchange ; main ; exchange on one stream.  (With one NCCL communicator per CONTEXT, as in round 1,
// this schedule hung once earlier solvers of the same process had used theirs; the process-wide
// communicator of lbm_nccl.cuh removed that.)
int lbm2p_run_slab(lbm2p_ctx *c, int nsteps, void *cuda_stream) {
    CTX2(c);
    if (!c->inited || !c->halo) FAIL2(c, -4, "needs an initialised x-slab context");
    if (!c->comm_ready) FAIL2(c, -4, "lbm2p_comm_init has not been called");
    if (nsteps <= 0) return 0;
    CU2(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    c->stream = st;
    if (!c->pipe_valid) {
        int r = stage2(c, 0, st);
        if (r < 0) return r;
        r = exchange2(c, 0, st);
        if (r) return r;
        nsteps -= 1;
    }
    const int nx = c->cfg.nx, own = nx - 2;
    static const bool want_overlap = getenv("LBM3D_2P_OVERLAP") ? atoi(getenv("LBM3D_2P_OVERLAP")) != 0 : true;
    if (!want_overlap || own < 3) {
        for (int it = 0; it < nsteps; ++it) {
            int r = stage2(c, 1, st);                  // rho_r, rho_b, psi of the step just collided
            if (r < 0) return r;
            r = exchange2(c, 1, st);                   // psi of the boundary planes -> neighbours' ghosts
            if (r) return r;
            r = stage2(c, 2, st);                      // stream/BC/macro + next collision
            if (r < 0) return r;
            r = exchange2(c, 0, st);                   // f* (5 + 5 populations) and colour records
            if (r) return r;
        }
        return 0;
    }
    bool xa_pending = false;
    static const int sync_every = getenv("LBM3D_2P_SYNC_EVERY") ? atoi(getenv("LBM3D_2P_SYNC_EVERY")) : 0;
    for (int it = 0; it < nsteps; ++it) {
        int r = 0;
        // optional bound on how far the host runs ahead of the two streams
        if (sync_every > 0 && it > 0 && it % sync_every == 0) CU2(c, cudaStreamSynchronize(c->comm_stream));
        if (!c->colour_valid) {
            r = launch_planes2(c, 0, 2, nx - 2, st);                       // colour, interior planes
            if (r) return r;
            if (xa_pending) CU2(c, cudaStreamWaitEvent(st, c->ev_xa, 0));  // ghost records have arrived
            xa_pending = false;
            r = launch_planes2(c, 0, 1, 2, st);                            // colour, boundary planes
            if (r) return r;
            r = launch_planes2(c, 0, nx - 2, nx - 1, st);
            if (r) return r;
            colour_done(c);
        } else if (xa_pending) {
            CU2(c, cudaStreamWaitEvent(st, c->ev_xa, 0));
            xa_pending = false;
        }
        CU2(c, cudaEventRecord(c->ev_cb, st));
        CU2(c, cudaStreamWaitEvent(c->comm_stream, c->ev_cb, 0));
        r = exchange2(c, 1, c->comm_stream);                               // XB: psi
        if (r) return r;
        CU2(c, cudaEventRecord(c->ev_xb, c->comm_stream));
        r = launch_planes2(c, 1, 2, nx - 2, st);                           // main, interior planes
        if (r) return r;
        CU2(c, cudaStreamWaitEvent(st, c->ev_xb, 0));                      // ghost psi has arrived
        r = launch_planes2(c, 1, 1, 2, st);                                // main, boundary planes
        if (r) return r;
        r = launch_planes2(c, 1, nx - 2, nx - 1, st);
        if (r) return r;
        c->cur ^= 1;
        c->colour_valid = false;
        c->macro_valid = false;
        c->F_valid = false;
        CU2(c, cudaEventRecord(c->ev_mb, st));
        CU2(c, cudaStreamWaitEvent(c->comm_stream, c->ev_mb, 0));
        r = exchange2(c, 0, c->comm_stream);                               // XA: f*, colour records
        if (r) return r;
        CU2(c, cudaEventRecord(c->ev_xa, c->comm_stream));
        xa_pending = true;
    }
    if (xa_pending) CU2(c, cudaStreamWaitEvent(st, c->ev_xa, 0));
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
#undef GETTER

// node-linear view of a colour field: dense storage holds it that way (solid nodes: 0); sparse
// storage scatters the compact array over a zero-filled (or, for psi, input-filled) lattice
static int get_colour_field(lbm2p_ctx *c, float *dst, const float *field, const float *solid_fill) {
    int r = sync2(c, false, false);
    if (r) return r;
    if (!c->sparse && solid_fill == nullptr) return out2(c, dst, field, c->N * sizeof(float));
    float *tmp = nullptr;
    CU2(c, cudaMalloc(&tmp, c->N * sizeof(float)));
    cudaError_t e = cudaSuccess;
    if (c->sparse) {
        e = solid_fill ? cudaMemcpyAsync(tmp, solid_fill, c->N * sizeof(float), cudaMemcpyDeviceToDevice, c->stream)
                       : cudaMemsetAsync(tmp, 0, c->N * sizeof(float), c->stream);
        if (e == cudaSuccess && c->nf) {
            k2p_scatter<<<nblocks(c->nf, 256), 256, 0, c->stream>>>(c->sp.d_lin, field, c->nf, tmp);
            e = cudaGetLastError();
        }
    } else {
        k2p_merge_psi<<<nblocks(c->N, 256), 256, 0, c->stream>>>(c->d_solid, solid_fill, field, tmp, c->N);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(dst, tmp, c->N * sizeof(float), cudaMemcpyDefault);
    cudaFree(tmp);
    CU2(c, e);
    c->launches++;
    return 0;
}

// rho_r (which = 0) or rho_b (1): one component of the interleaved array, as its own field
static int get_colour_component(lbm2p_ctx *c, float *dst, int which) {
    int r = sync2(c, false, false);
    if (r) return r;
    const size_t n = c->sparse ? c->sp.stride : c->N;
    float *tmp = nullptr;
    CU2(c, cudaMalloc(&tmp, (n ? n : 1) * sizeof(float)));
    k2p_split<<<nblocks(n ? n : 1, 256), 256, 0, c->stream>>>(c->d_rrb[c->rcur], which, n, tmp);
    cudaError_t e = cudaGetLastError();
    c->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(tmp); CU2(c, e); }
    r = get_colour_field(c, dst, tmp, nullptr);
    cudaFree(tmp);
    return r;
}

int lbm2p_get_rho_r(lbm2p_ctx *c, float *dst) {
    CTX2(c);
    if (!dst) FAIL2(c, -1, "null destination");
    return get_colour_component(c, dst, 0);
}

int lbm2p_get_rho_b(lbm2p_ctx *c, float *dst) {
    CTX2(c);
    if (!dst) FAIL2(c, -1, "null destination");
    return get_colour_component(c, dst, 1);
}

int lbm2p_get_psi(lbm2p_ctx *c, float *dst) {
    CTX2(c);
    if (!dst) FAIL2(c, -1, "null destination");
    return get_colour_field(c, dst, c->d_psi, c->d_psi0);      // solid nodes keep the input value
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
    if (c->sparse) {
        float *tmp = nullptr;
        CU2(c, cudaMalloc(&tmp, c->N * sizeof(float)));
        float *tmp2 = nullptr;
        cudaError_t e = cudaMalloc(&tmp2, c->N * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(tmp, psi, c->N * sizeof(float), cudaMemcpyDefault);
        if (e == cudaSuccess && c->nf) {
            k2p_gather<<<nblocks(c->nf, 256), 256>>>(c->sp.d_lin, tmp, c->nf, c->d_psi);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(tmp, rho_r, c->N * sizeof(float), cudaMemcpyDefault);
        if (e == cudaSuccess) e = cudaMemcpy(tmp2, rho_b, c->N * sizeof(float), cudaMemcpyDefault);
        if (e == cudaSuccess && c->nf) {
            k2p_join<<<nblocks(c->nf, 256), 256>>>(tmp, tmp2, c->sp.d_lin, c->nf, c->d_rrb[c->rcur]);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(tmp);
        cudaFree(tmp2);
        CU2(c, e);
    } else {
    CU2(c, cudaMemcpy(c->d_psi, psi, c->N * sizeof(float), cudaMemcpyDefault));
    {
        float *t0 = nullptr, *t1 = nullptr;
        cudaError_t e = cudaMalloc(&t0, c->N * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&t1, c->N * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(t0, rho_r, c->N * sizeof(float), cudaMemcpyDefault);
        if (e == cudaSuccess) e = cudaMemcpy(t1, rho_b, c->N * sizeof(float), cudaMemcpyDefault);
        if (e == cudaSuccess) {
            k2p_join<<<nblocks(c->N, 256), 256>>>(t0, t1, nullptr, c->N, c->d_rrb[c->rcur]);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(t0);
        cudaFree(t1);
        CU2(c, e);
    }
    k2p_bake_psi<<<nblocks(c->N, 256), 256>>>(c->d_solid, c->d_psi, (float)c->psi_solid, c->N);
    }
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
    CU2(c, max_v_reduce(c->d_v, c->N, c->d_scalar, c->stream, out));
    c->launches++;
    return 0;
}

}  // extern "C"
