/* lbm3d_2phase.h -- C ABI of the B200-native two-phase (colour-gradient) D3Q19 MRT step.
 *
 * Replaces the kernels of the reference script 2phase/lbm_solver_3d_2phase.py (and its
 * pointer-sparse twin lbm_solver_3d_2phase_sparse.py, which differs only in allocation):
 *   init :173, init_geo :194, static_init :205, colission :302 (Compute_C :259,
 *   Compute_S_local :278, GuoF :241), streaming1 :431, Boundary_condition_psi :445,
 *   Boundary_condition :491, streaming3 :587, main loop :626-632.
 * The reference has no class for this solver (its README lists "wrap functions into class" as
 * to-do); taichi_lbm3d_b200/lbm_solver_3d_2phase.py provides one over this ABI with the
 * script's global names as attributes.
 *
 * Conventions are those of lbm3d.h (status codes, borrowed pointers, host-or-device
 * destinations, one context per GPU, fp32 everywhere, arrays [nx][ny][nz](,C) C order).
 */
#ifndef LBM3D_2PHASE_H
#define LBM3D_2PHASE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbm2p_ctx lbm2p_ctx;

/* lbm2p_config.reserved: x-slab flags (multi-GPU, new work: the reference is single-device) */
#define LBM2P_HALO_X 1     /* planes x=0 and x=nx-1 are ghost planes owned by the slab neighbours */
#define LBM2P_HOLDS_X0 2   /* this slab's first owned plane is the global x0 face */
#define LBM2P_HOLDS_X1 4   /* this slab's last owned plane is the global x1 face */
#define LBM2P_SPARSE 8     /* compact fluid-node list + pull table instead of the full lattice: the
                              storage of 2phase/lbm_solver_3d_2phase_sparse.py (:54-75), as the
                              single-phase sparse mode does it; combines with LBM2P_HALO_X */

typedef struct {
    int32_t nx, ny, nz;   /* :17 */
    int32_t strict;       /* 1: oracle evaluation order, no FMA contraction (verification) */
    int32_t device;
    int32_t reserved;     /* bit set of LBM2P_SPARSE, LBM2P_HALO_X, LBM2P_HOLDS_X0, LBM2P_HOLDS_X1 */
} lbm2p_config;

int lbm2p_create(const lbm2p_config *cfg, lbm2p_ctx **out);
int lbm2p_destroy(lbm2p_ctx *ctx);
const char *lbm2p_last_error(const lbm2p_ctx *ctx);

/* ---- parameters: the script's module-level globals :16-39 (all before lbm2p_init) ------- */
int lbm2p_set_geometry(lbm2p_ctx *ctx, const int8_t *solid_host_or_dev);      /* solid.from_numpy :615 */
int lbm2p_set_phase(lbm2p_ctx *ctx, const float *psi_host_or_dev);            /* psi.from_numpy :616 */
/* niu_l, niu_g, psi_solid, CapA (:21-24); the relaxation constants wl, wg, lg0, l1, l2, g1, g2
 * of :100-108 are evaluated in double exactly as the Python source does */
int lbm2p_set_fluid(lbm2p_ctx *ctx, double niu_l, double niu_g, double psi_solid, double CapA);
int lbm2p_set_force(lbm2p_ctx *ctx, const float force[3]);                     /* fx,fy,fz :19 */
/* flow faces (:27-32): type 0 periodic, 1 fixed pressure (rho), 2 the script's velocity form
 * (:500-504; its bc_vel fields are never written, so vel should be zero to match it) */
int lbm2p_set_bc(lbm2p_ctx *ctx, int face, int type, float rho, const float vel[3]);
/* phase-field faces (:34-39): type 0 periodic, 1 constant psi */
int lbm2p_set_psi_bc(lbm2p_ctx *ctx, int face, int type, float psi);
int lbm2p_set_inverse_matrix(lbm2p_ctx *ctx, const float invM[361]);          /* verification mode */

/* static_init + init (:205-228, :173-186) */
int lbm2p_init(lbm2p_ctx *ctx);
/* nsteps iterations of the main loop body :626-632; two kernel launches per step */
int lbm2p_step(lbm2p_ctx *ctx, int nsteps, void *cuda_stream);
int64_t lbm2p_launch_count(const lbm2p_ctx *ctx);
int lbm2p_synchronize(lbm2p_ctx *ctx);

/* fields as the script's to_numpy() shows them after an iteration */
int lbm2p_get_rho(lbm2p_ctx *ctx, float *dst);      /* [nx][ny][nz]     */
int lbm2p_get_v(lbm2p_ctx *ctx, float *dst);        /* [nx][ny][nz][3]  */
int lbm2p_get_F(lbm2p_ctx *ctx, float *dst);        /* [nx][ny][nz][19] */
int lbm2p_get_psi(lbm2p_ctx *ctx, float *dst);      /* [nx][ny][nz]; solid nodes keep the input value */
int lbm2p_get_rho_r(lbm2p_ctx *ctx, float *dst);
int lbm2p_get_rho_b(lbm2p_ctx *ctx, float *dst);
int lbm2p_get_solid(lbm2p_ctx *ctx, int8_t *dst);
/* replace the whole state (restart / perturbed start): F [..][19], rho, v [..][3], psi, rho_r, rho_b */
int lbm2p_set_state(lbm2p_ctx *ctx, const float *F, const float *rho, const float *v, const float *psi,
                    const float *rho_r, const float *rho_b);
int lbm2p_get_max_v(lbm2p_ctx *ctx, float *out);

/* ---- multi-GPU x-slabs (contexts created with LBM2P_HALO_X) -------------------------------
 * The collision of a node needs psi of its 18 neighbours and the colour pass the records of its
 * 18 pull sources, so a step has TWO exchanges across every cut:
 *   stage 0 (after the main pass):   the 5 populations of f* that cross the cut + the part of
 *                                    the colour record the collision writes (v, q and the
 *                                    interface vector) of the boundary plane, 13 floats per node
 *   stage 1 (after the colour pass): psi, rho_r, rho_b of the boundary plane, 3 floats per node
 * A plane of a dense slab has ny*nz nodes, a plane of a sparse slab its fluid nodes, so the four
 * halo planes (0 = left ghost, 1 = first owned, 2 = last owned, 3 = right ghost) differ in size:
 * lbm2p_halo_count(plane) nodes; lbm2p_halo_floats(stage) is the largest message of that stage.
 * lbm2p_run_slab drives colour ; exchange(1) ; main ; exchange(0) with ncclSend/ncclRecv.  The
 * pack / unpack / stage entry points let another transport (torch.distributed, or several
 * emulated ranks in one process) run the same schedule:
 *   stage(0) ; exchange(0) ; repeat { stage(1) ; exchange(1) ; stage(2) ; exchange(0) }
 * side 0 = towards x-1 (packs local plane 1, fills ghost plane 0), side 1 = towards x+1. */
int64_t lbm2p_halo_floats(lbm2p_ctx *ctx, int stage);
int64_t lbm2p_halo_count(lbm2p_ctx *ctx, int plane);
int lbm2p_halo_pack(lbm2p_ctx *ctx, int stage, int side, float *dst_dev, void *cuda_stream);
int lbm2p_halo_unpack(lbm2p_ctx *ctx, int stage, int side, const float *src_dev, void *cuda_stream);
/* stage 0: first collision of the user-visible state (returns 1 if already done), 1: colour
 * pass, 2: main pass */
int lbm2p_slab_stage(lbm2p_ctx *ctx, int stage, void *cuda_stream);
int lbm2p_comm_unique_id(void *out128);
int lbm2p_comm_init(lbm2p_ctx *ctx, const void *id128, int world, int rank);
int lbm2p_run_slab(lbm2p_ctx *ctx, int nsteps, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LBM3D_2PHASE_H */
