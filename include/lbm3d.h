/* lbm3d.h -- C ABI of the B200-native D3Q19 MRT lattice-Boltzmann time step.
 *
 * This is the drop-in boundary for the hot path of yjhp1016/taichi_LBM3D.  The
 * reference has no FFI: its "kernels" are the @ti.kernel methods of
 *   Single_phase/LBM_3D_SinglePhase_Solver.py  (class LB3D_Solver_Single_Phase)
 * called from step() (:477-481).  Each entry point below names the reference method
 * it replaces (file:line relative to that file unless another file is given).  The
 * Python class in taichi_lbm3d_b200/LBM_3D_SinglePhase_Solver.py binds these through
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - every function returns 0 on success or a negative lbm_status; nothing throws
 *    across the ABI; lbm_last_error(ctx) gives the message of the last failure.
 *  - a context owns its device buffers on ONE GPU; one context per GPU; a context is
 *    not thread-safe.  Pointers passed in are borrowed for the duration of the call.
 *  - "host_or_dev" pointers may be host or device memory (cudaMemcpyDefault).
 *  - dense user-visible arrays use the layout the reference's to_numpy() exposes:
 *    C order [nx][ny][nz] (z fastest), vectors with a trailing component axis.
 *  - all arithmetic is fp32 (ti.f32 in the reference, :32-35); geometry is int8 (:57).
 *  - work is enqueued on the stream given to lbm_step; getters synchronise it.
 */
#ifndef LBM3D_H
#define LBM3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM3D_ABI_VERSION 1

typedef struct lbm_ctx lbm_ctx;

typedef enum {
    LBM_OK = 0,
    LBM_ERR_INVALID = -1,   /* bad argument / call order */
    LBM_ERR_CUDA = -2,      /* CUDA runtime error (message in lbm_last_error) */
    LBM_ERR_NOMEM = -3,
    LBM_ERR_STATE = -4      /* e.g. step before init */
} lbm_status;

/* faces: order of the reference's Boundary_condition kernel (:272-370); on shared
 * edges/corners the LATER face wins because the reference runs the face loops in
 * this order. */
enum { LBM_FACE_X0 = 0, LBM_FACE_X1, LBM_FACE_Y0, LBM_FACE_Y1, LBM_FACE_Z0, LBM_FACE_Z1 };
/* :22  "0=periodic, 1= fix pressure, 2=fix velocity" */
enum { LBM_BC_PERIODIC = 0, LBM_BC_PRESSURE = 1, LBM_BC_VELOCITY = 2 };

typedef struct {
    int32_t nx, ny, nz;   /* lattice extents of this context (ctor :11) */
    int32_t sparse;       /* 0: direct-addressed dense lattice (:31-35);
                             1: compacted fluid-node list + 18-neighbour table, replaces the
                                pointer/dense SNode tree (:36-44); two population buffers (A-B);
                                at most 2^30 fluid nodes per context (split larger domains
                                into x-slabs);
                             2: the same list stepped IN PLACE on one buffer (AA pattern): odd
                                steps pull through the table and store back into the pulled
                                locations, even steps are purely local -- half the memory and
                                the table is read every other step.  Falls back to 1 with
                                halo_x (slab exchange needs both buffers);
                             3: the dense lattice of mode 0 stepped IN PLACE the same way (half
                                the memory: 1024^3 fits one 180 GB B200).  Falls back to 0 with
                                halo_x */
    int32_t strict;       /* 0: factored MRT transform, FMA allowed (production);
                             1: oracle evaluation order, no FMA contraction -> bit-identical
                                to oracle/ref_single_phase.c (verification mode) */
    int32_t halo_x;       /* 0: periodic_index wraps x (:250-251);
                             1: planes x=0 and x=nx-1 are ghost planes owned by the x-slab
                                neighbours (multi-GPU); they are read, never updated */
    int32_t device;       /* CUDA device ordinal */
    int32_t x_face_mask;  /* halo_x=1 only: bit0 = this slab holds the global x0 face (its first owned
                             plane), bit1 = it holds the global x1 face (its last owned plane) */
} lbm_config;

/* ---- lifetime ---------------------------------------------------------------------- */
int lbm_abi_version(void);
/* replaces LB3D_Solver_Single_Phase.__init__ (:11-114): allocates solid, rho, v, populations */
int lbm_create(const lbm_config *cfg, lbm_ctx **out);
int lbm_destroy(lbm_ctx *ctx);
const char *lbm_last_error(const lbm_ctx *ctx);   /* ctx may be NULL: last create error */
/* lbm_destroy / a repeated lbm_init keep the device buffers of a context (at most LBM3D_POOL_MB
 * MiB in total, default 8192, 0 = never) for the next context of this process that asks for the
 * same sizes, which then makes no cudaMalloc / cudaFree call (csrc/lbm_devpool.cuh).  This gives the cached memory back to the driver; returns the bytes freed.
 * (No counterpart in the reference: Taichi's runtime owns its device memory pool.) */
long long lbm_pool_trim(void);

/* ---- parameters (setters :405-458; all must precede lbm_init) ------------------------- */
/* solid.from_numpy / init_geo (:173-177): int8 [nx][ny][nz], >0 = solid */
int lbm_set_geometry(lbm_ctx *ctx, const int8_t *solid_host_or_dev);
/* set_bc_vel_* / set_bc_rho_* (:405-451): type 2 uses vel (rho ignored), type 1 uses rho */
int lbm_set_bc(lbm_ctx *ctx, int face, int type, float rho, const float vel[3]);
/* form of the fixed-velocity faces: 0 = the class (:283-288), F = feq(1, u) on all 19 populations;
 * 1 = the script copies of the solver (Single_phase/lbm_solver_3d.py:253,268), in place for
 * s = 0..18:  F[s] = feq(LR[s], 1, u) - F[LR[s]] + feq(s, 1, u).  Before lbm_init. */
int lbm_set_vel_bc_form(lbm_ctx *ctx, int script_form);
/* Grey-scale lattice (replaces ns.from_numpy of Grey_Scale/lbm_solver_3d_Macro_Sukop.py:343 and
 * its streaming0 / streaming1, :233-247): float [nx][ny][nz], the solid fraction of every node.
 * Post-collision populations are blended with the opposite population of the node ahead,
 * f2[i][s] = f[i][s] + ns[i] (f[i+e_s][LR[s]] - f[i][s]), and streamed to every neighbour without
 * bounce-back; links that leave a solid node (lbm_set_geometry; the script uses solid = int(ns))
 * deliver the rest value w[s], as in the script, which never writes them.  With
 * lbm_set_relaxation(3 niu + 1/2 ...), lbm_set_guo_form(1) this is that script's time step.
 * Dense two-buffer storage on one GPU, periodic and fixed-pressure faces; NULL switches it off.
 * Before lbm_init. */
int lbm_set_grey_scale(lbm_ctx *ctx, const float *ns_host_or_dev);
/* set_force (:457); force_flag as :137-140 */
int lbm_set_force(lbm_ctx *ctx, const float force[3]);
/* form of the Guo force term in the collision: 0 = the class (:236, parts divided by 3 and 9:
 * effective body force f/9), 1 = un-scaled, as in the other copy of the solver
 * (Phase_change/LBM_3D_SinglePhase_Solver.py:235) and the script variants.  Before lbm_init. */
int lbm_set_guo_form(lbm_ctx *ctx, int unscaled);
/* per-node force [nx][ny][nz][3] (host or device; copied): the array form of the reference's
 * override point cal_local_force(i,j,k) (:217-220, overridden by
 * Phase_change/LBM_3D_SinglePhase_Solute_Solver.py:185-190 to add buoyancy).  Used by the
 * collision (:230-238) and by the velocity shift of streaming3 (:388) in place of the uniform
 * force; after lbm_init, may be replaced between steps; NULL returns to lbm_set_force. */
int lbm_set_force_field(lbm_ctx *ctx, const float *force3_host_or_dev);
/* set_viscosity + the relaxation rates of init_simulation (:126-131), evaluated in
 * double exactly as the Python source does, rounded once to fp32.
 * textbook_tau = 0: tau = niu/3 + 0.5 (the class, :127); 1: tau = 3 niu + 0.5 (:126,
 * the form every other copy of the solver uses). */
int lbm_set_viscosity(lbm_ctx *ctx, double niu, int textbook_tau);
/* alternatively give the 19 diagonal rates S_dig (:131) directly */
int lbm_set_relaxation(lbm_ctx *ctx, const float S[19]);

/* ---- init_simulation (:118-149): static_init + init (:160-170) ------------------------- */
/* builds the link flags (dense) or the compacted fluid list + neighbour table (sparse)
 * from the geometry, and sets rho=1, v=0, f=F=w. */
int lbm_init(lbm_ctx *ctx);

/* ---- step (:477-481): colission -> streaming1 -> Boundary_condition -> streaming3 ------- */
/* nsteps reference steps, fused: one kernel launch per step on `cuda_stream`
 * (a cudaStream_t, NULL = default stream).  Asynchronous. */
int lbm_step(lbm_ctx *ctx, int nsteps, void *cuda_stream);
/* number of kernels the context has launched so far (bench evidence) */
int64_t lbm_launch_count(const lbm_ctx *ctx);
/* wait for all enqueued work */
int lbm_synchronize(lbm_ctx *ctx);

/* ---- fields (what rho.to_numpy(), v.to_numpy(), F.to_numpy() return) ------------------ */
/* each getter first brings the user-visible state up to date (the streaming3 pass
 * :372-392 of the last step) and synchronises. */
int lbm_get_rho(lbm_ctx *ctx, float *dst_host_or_dev);              /* [nx][ny][nz]     */
int lbm_get_v(lbm_ctx *ctx, float *dst_host_or_dev);                /* [nx][ny][nz][3]  */
int lbm_get_F(lbm_ctx *ctx, float *dst_host_or_dev);                /* [nx][ny][nz][19] */
int lbm_get_solid(lbm_ctx *ctx, int8_t *dst_host_or_dev);           /* [nx][ny][nz]     */
/* F.from_numpy / rho.from_numpy / v.from_numpy: overwrite one field of the state */
int lbm_set_rho(lbm_ctx *ctx, const float *src_host_or_dev);
int lbm_set_v(lbm_ctx *ctx, const float *src_host_or_dev);
int lbm_set_F(lbm_ctx *ctx, const float *src_host_or_dev);
/* get_max_v + cal_max_v (:394-402) */
int lbm_get_max_v(lbm_ctx *ctx, float *out);
/* probe: the fields at n selected nodes (linear index i*ny*nz + j*nz + k), what F[i,j,k], rho[i,j,k],
 * v[i,j,k] read in the reference -- without copying whole lattices to the host (F of a 512^3 lattice
 * is 10 GB).  Any of the three outputs may be NULL; F_out is [n][19], v_out [n][3]. */
int lbm_get_nodes(lbm_ctx *ctx, int64_t n, const int64_t *index_host_or_dev, float *F_out_host_or_dev,
                  float *rho_out_host_or_dev, float *v_out_host_or_dev);

/* ---- sparse storage tables (bit-exact compaction tests; north_star item 1) ------------ */
int lbm_get_num_fluid(lbm_ctx *ctx, int64_t *n_fluid);
/* linear index i*ny*nz + j*nz + k of every stored fluid node, ascending; [n_fluid] */
int lbm_get_fluid_index(lbm_ctx *ctx, int64_t *dst_host_or_dev);
/* pull table [18][n_fluid]: row s-1 holds, for direction s=1..18, the compact index of
 * the node i - e_s (periodic_index wrapped, :247-257) or -1 when that node is solid
 * (half-way bounce-back, :267-268) */
int lbm_get_neighbor_table(lbm_ctx *ctx, int32_t *dst_host_or_dev);
/* dense mode: per-node link word [nx][ny][nz]; bit s (1..18) set = pull source of
 * direction s is solid; bit 19 = node solid; bits 20-22 = winning BC face + 1;
 * bit 23 = pressure BC uses the (zero) velocity of a solid inward neighbour (:278) */
int lbm_get_link_flags(lbm_ctx *ctx, uint32_t *dst_host_or_dev);

/* ---- multi-GPU x-slabs (halo_x = 1) -----------------------------------------------------
 * The reference is single-device; this is the new decomposition.  A context
 * then holds one x-slab plus one ghost plane on each side.  The five populations with
 * e_x = +1 (s = 1,7,9,11,13) leave through the right face and the five with e_x = -1
 * (s = 2,8,10,12,14) through the left face (:183-187); nothing else crosses a cut.
 *
 * plane ids: 0 = left ghost (x=0), 1 = first owned (x=1), 2 = last owned (x=nx-2),
 *            3 = right ghost (x=nx-1).  lbm_halo_count = stored nodes in that plane
 *            (dense: ny*nz; sparse: its fluid nodes).
 * lbm_halo_pack(side)   side 0: e_x=-1 populations of plane 1 -> dst (goes to the LEFT rank)
 *                       side 1: e_x=+1 populations of plane 2 -> dst (goes to the RIGHT rank)
 *                       dst holds 5*count floats, [5][count].
 * lbm_halo_unpack(side) side 0: src (sent by the LEFT rank's pack(1)) -> e_x=+1 of plane 0
 *                       side 1: src (sent by the RIGHT rank's pack(0)) -> e_x=-1 of plane 3
 * `which` = 0 acts on the CURRENT post-collision buffer, 1 on the NEXT one (the output of
 * lbm_step_planes before lbm_step_flip).  Plain sequence:
 *   lbm_step_begin ; exchange(0) ; repeat { lbm_step(1) ; exchange(0) }
 * Overlapped sequence per step: lbm_step_planes(first owned), lbm_step_planes(last owned) ;
 * exchange(1) on a side stream || lbm_step_planes(interior) ; join ; lbm_step_flip. */
int64_t lbm_halo_count(lbm_ctx *ctx, int plane);
int lbm_halo_pack(lbm_ctx *ctx, int side, int which, float *dst_dev, void *cuda_stream);
int lbm_halo_unpack(lbm_ctx *ctx, int side, int which, const float *src_dev, void *cuda_stream);
/* The whole slab loop in native code: NCCL is bound at run time (dlopen of the libnccl.so.2
 * already loaded by torch).  Rank 0 calls lbm_comm_unique_id (128 bytes), the host side
 * broadcasts it (torch.distributed), every rank calls lbm_comm_init after lbm_init.
 * lbm_run_slab = nsteps x { boundary planes ; ncclSend/ncclRecv of the 5+5 face populations
 * on a side stream || interior planes ; join } when overlap != 0.  world = 1 needs no NCCL
 * (the ring closes on the slab itself). */
/* Direct peer-memory halo (dense storage, one process per GPU on one NVLink / NVSwitch node): the
 * boundary-plane kernel of lbm_run_slab stores the five crossing populations straight into the
 * neighbours' ghost planes and the ranks order themselves with flags in device memory, so a step
 * has no pack, no ncclSend/ncclRecv and no unpack.  After lbm_comm_init: every rank exports a
 * 256-byte blob (CUDA IPC handles of its two population buffers and its flag words + its layout),
 * the host side exchanges the blobs, every rank connects to its left and right neighbour's, and
 * -- only when EVERY rank connected -- all enable it; otherwise all keep the NCCL exchange. */
#define LBM_P2P_BLOB_BYTES 256
int lbm_p2p_export(lbm_ctx *ctx, void *blob256);
int lbm_p2p_connect(lbm_ctx *ctx, const void *left_blob256, const void *right_blob256);
int lbm_p2p_enable(lbm_ctx *ctx, int on);
/* unmap the neighbours' buffers; every rank calls it (then a barrier) BEFORE any rank destroys its
 * context: exported memory must outlive its mappings */
int lbm_p2p_disconnect(lbm_ctx *ctx);
/* 1 when this process already holds the NCCL communicator for (world, rank) on the current device
 * (one per process, shared by all contexts): lbm_comm_init then needs no unique id */
int lbm_comm_ready(int world, int rank);
int lbm_comm_unique_id(void *out128);
int lbm_comm_init(lbm_ctx *ctx, const void *id128, int world, int rank);
int lbm_run_slab(lbm_ctx *ctx, int nsteps, int overlap, void *cuda_stream);
/* first collision of the user-visible state (:222-241) if the pipeline is not running
 * yet; counts as the collision half of the next step.  Returns 1 if already running. */
int lbm_step_begin(lbm_ctx *ctx, void *cuda_stream);
int lbm_step_planes(lbm_ctx *ctx, int x_begin, int x_end, void *cuda_stream);
int lbm_step_flip(lbm_ctx *ctx);

/* verification mode only: the 19x19 inv_M (:83,:110) as the caller's np.linalg.inv
 * produced it (LAPACK leaves 1e-17 noise in the structural zeros); default = exact. */
int lbm_set_inverse_matrix(lbm_ctx *ctx, const float invM[361]);

/* raw device pointers for zero-copy wrapping (torch.as_tensor via __cuda_array_interface__) */
enum { LBM_BUF_F_CUR = 0, LBM_BUF_F_NEXT = 1, LBM_BUF_RHO = 2, LBM_BUF_V = 3, LBM_BUF_FLAGS = 4 };
int lbm_get_device_ptr(lbm_ctx *ctx, int which, void **ptr, size_t *bytes);
/* layout of LBM_BUF_F_*: out[0] = 0 SoA planes [19][stride] | 1 row-blocked [row][19][nzp]
 * (row = i*ny+j, nzp = nz rounded up to 32) | 2 compact fluid list [19][stride];
 * out[1] = elements between planes s and s+1; out[2] = elements between z-rows (dense);
 * out[3] = elements per buffer */
int lbm_get_layout(lbm_ctx *ctx, int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* LBM3D_H */
