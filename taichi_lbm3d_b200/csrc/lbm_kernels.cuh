// Interface between the C-ABI layer (lbm_api.cu) and the kernel translation unit
// (lbm_kernels.cu, compiled twice: production arithmetic -> namespace lbm_fast,
// -DLBM_STRICT -fmad=false -> namespace lbm_strict).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lbm_d3q19.cuh"

// ---- link word of the dense lattice (lbm_get_link_flags) ---------------------------------
#define FL_LINK_MASK 0x0007FFFEu  /* bits 1..18: pull source of direction s is solid        */
#define FL_SOLID (1u << 19)       /* node is solid                                           */
#define FL_BC_SHIFT 20            /* bits 20..22: winning BC face + 1 (0 = not a BC node)    */
#define FL_BC_MASK 7u
#define FL_PIN_SOLID (1u << 23)   /* pressure BC: inward neighbour is solid -> v = 0 (:278)  */
#define FL_AT_X0 (1u << 24)       /* node sits on a face across which periodic_index wraps   */
#define FL_AT_X1 (1u << 25)
#define FL_AT_Y0 (1u << 26)
#define FL_AT_Y1 (1u << 27)
#define FL_AT_Z0 (1u << 28)
#define FL_AT_Z1 (1u << 29)

#define FL_NEAR_SOLID (1u << 30)  /* two-phase: a node of the phase-field stencil is solid         */
#define FL_EXCEPTION (1u << 31)   /* sparse: pull sources listed explicitly in the exception table */

enum { MODE_STEP = 0, MODE_EXTRACT = 1, MODE_COLLIDE = 2 };
enum { AA_OFF = 0, AA_ODD = 1, AA_EVEN = 2 };
// dense node classes: BULK = fluid, no solid link / wrap / BC (the speculative pull is final);
// SOLID = nothing to do; SPECIAL = fluid that needs its link word; SOLID_WRITE = solid node
// sharing a 32-byte sector with a fluid node: it stores (dead) values so that the sector is
// written whole and L2 never has to read-modify-write it
enum { NODE_BULK = 0, NODE_SOLID = 1, NODE_SPECIAL = 2, NODE_SOLID_WRITE = 3 };

struct StepArgs {
    // populations, SoA, one plane per direction (post-collision state of the pipeline).
    // Plane base pointers are passed ready-made so that every access is
    // "uniform 64-bit base + 32-bit node index" (one IMAD.WIDE per load/store):
    //   pown[s]  plane s of the input buffer          pout[s]  plane s of the output buffer
    //   ppull[s] = pown[s] - (ex*ny*nz + ey*nz + ez)  pull source of a node without wrap
    const float *pown[19];
    const float *ppull[19];
    float *pout[19];
    size_t stride;
    // sparse: node range processed by this launch: [first, first+count)
    uint32_t first, count;
    // dense: z-rows [row_first, row_first+row_count) (row = i*ny + j); element of node (row, k)
    // in a population plane is row*prow + k.  SoA layout: prow = nz, planes `stride` apart;
    // row-blocked layout [row][19][nzp]: prow = 19*nzp, planes nzp apart.
    uint32_t row_first, row_count, prow;
    // one launch over TWO disjoint ranges (the two boundary planes of an x-slab): dense, launch row
    // r >= row_split maps to row_first + r + row_skip; sparse, blocks >= nb1 walk [first2, first2+count2)
    uint32_t row_split, row_skip;
    uint32_t nb1, first2, count2;
    int spec;                       // 1: speculative pull (few solid nodes)
    int nx, ny, nz;                 // extents of this context's lattice (incl. ghost planes)
    int halo_x;                     // 1: planes 0 and nx-1 are ghost planes, x never wraps
    // dense: link word per node.  sparse: BC word per stored node (bits 20..23), may be null
    const uint32_t *flags;
    // dense: node class byte (NODE_BULK / NODE_SOLID / NODE_SPECIAL); only NODE_SPECIAL nodes
    // (solid links, periodic wrap, face BC) read their 32-bit link word
    const uint8_t *cls;
    // sparse only
    const int32_t *nbr[18];         // rows of the full [18][stride] pull table, -1 = bounce
    // compressed pull table (default): link word per node (`flags`: bits 1..18 source solid,
    // 20..23 BC, 31 exception) + rb[k][i] = fluid rank of the position (x-ex, y-ey, z) in the
    // k-th neighbour z-row, k over (ex,ey) = (1,0),(-1,0),(0,1),(0,-1),(1,1),(-1,-1),(1,-1),(-1,1).
    // The list is sorted with z fastest, so the three sources (z-1, z, z+1) of a neighbour row
    // are rb-1, rb, rb+fluid(center).  What is stored is how far that rank runs ahead of the node's
    // own index, rb[k][i] - i, as an 8-bit offset from its minimum over the 256-node block:
    //     rb[k][i] = blk[B][k] + i + rb8[k][i],  B = i / 256
    // (neighbouring z-rows fill at nearly the same rate, so over a block the lead changes by a few
    // tens): 8 x 1 B + 4 B = 12 B per node instead of 18 x 4 = 72 B.
    // Nodes for which the rank rule fails (periodic z wrap), and all nodes of a block in which a
    // lead varies by more than 255 (a periodic x / y wrap inside the block; < 1 % of the blocks of
    // a 512^3 sphere pack), carry FL_EXCEPTION and keep their 18 sources in exc[s-1][slot],
    // slot = blk[B][8] + rb8[0][i].
    const uint8_t *rb8[8];
    const int32_t *blk;             // [n_blocks][16]: 8 lead bases, exception-slot base, padding
    const int32_t *exc[18];
    int compressed;
    uint32_t prefetch_dist;         // nodes ahead whose table lines are pulled into L2 (0 = off)
    // sparse in-place (AA-pattern) stepping on ONE buffer (pown == pout planes):
    //   AA_OFF   two buffers, pull from pown, store to pout
    //   AA_ODD   buffer in natural layout (slot s of node i = f*_s(i)): pull like AA_OFF, then
    //            store f*_{LR[s]} back into the very location direction s was pulled from, which
    //            leaves every slot holding the population that ARRIVES there next step
    //   AA_EVEN  buffer in arrival layout: F_q(i) = slot LR[q] of node i, purely local, no
    //            table; stores f*_q into slot q (natural layout again)
    int aa;
    const uint32_t *lin;            // [n_fluid] linear index of each stored node
    // user-visible dense arrays (reference layout), used by MODE_EXTRACT / MODE_COLLIDE
    float *rho;                     // [N]
    float *v;                       // [N][3]
    float *F;                       // [N][19] or null
    // previous-step velocity of pressure-BC nodes (:279-281 read self.v before streaming3)
    float *vbc;                     // [sum of face sizes][3]
    uint32_t vbc_off[6];            // first slot of each face
    int force;                      // force_flag :137-140; 2 = per-node force array `ff`; 3 = `ff` + `ffm`
    // per-node force (replaces the cal_local_force override point :217-220): three planes in the
    // order of the stored nodes (dense: [3][N], sparse: [3][stride]); null = uniform P.force.
    // A fused launch closes step k (macro, :385-388) and opens step k+1 (collision, :230-238): when
    // the array was replaced in between, `ffm` is the one step k ran with (force == 3).
    const float *ff[3];
    const float *ffm[3];
    int has_bc;                     // any face with type != 0
    // grey-scale lattice (Grey_Scale/lbm_solver_3d_Macro_Sukop.py): solid fraction per node, dense
    // node order [N]; null = ordinary half-way bounce-back streaming
    const float *ns;
    // peer-memory halo of a dense x-slab (k_dense_peer, boundary-plane launch only; null otherwise):
    // peer_out[side][q] = plane of the q-th crossing population in the side's neighbour's OUTPUT
    // buffer, peer_delta[side] = element offset from this slab's boundary plane to that neighbour's
    // ghost plane, p2p = this rank's flag words, peer_flag[side] = the word of the neighbour that
    // counts this rank's completed launches, p2p_launch = index of this launch
    float *peer_out[2][5];
    long long peer_delta[2];
    int *p2p;
    int *peer_flag[2];
    int p2p_launch;
    d3q19::LbmParams P;
};

// ---- compressed sparse pull table: source of direction (EX,EY,EZ) for stored node i --------
__host__ __device__ constexpr int comp_row_slot(int ex, int ey) {
    return ey == 0 ? (ex > 0 ? 0 : 1) : (ex == 0 ? (ey > 0 ? 2 : 3) : (ex > 0 ? (ey > 0 ? 4 : 6) : (ey < 0 ? 5 : 7)));
}
// direction index of (ex,ey,0) for the four axis rows: its link bit tells whether the row's
// centre node (x-ex, y-ey, z) is solid
__host__ __device__ constexpr int comp_center_dir(int ex, int ey) {
    return ey == 0 ? (ex > 0 ? 1 : 2) : (ey > 0 ? 3 : 4);
}
template <int EX, int EY, int EZ>
__host__ __device__ __forceinline__ int32_t comp_source(uint32_t i, uint32_t mask, const int32_t (&rb)[8]) {
    if (EX == 0 && EY == 0) return EZ > 0 ? (int32_t)i - 1 : (int32_t)i + 1;
    const int32_t base = rb[comp_row_slot(EX, EY)];
    if (EZ == 0) return base;
    if (EZ > 0) return base - 1;                       // source at z-1: last fluid node before (.., z)
    return base + (((mask >> comp_center_dir(EX, EY)) & 1u) ? 0 : 1);   // source at z+1
}

#define LBM_DECLARE_KERNEL_API(NS)                                                            \
    namespace NS {                                                                            \
    cudaError_t launch_dense(int mode, const StepArgs &a, int block, cudaStream_t st);        \
    cudaError_t launch_sparse(int mode, const StepArgs &a, int block, cudaStream_t st);       \
    cudaError_t set_inverse_matrix(const float *invM361);                                     \
    }
LBM_DECLARE_KERNEL_API(lbm_fast)
LBM_DECLARE_KERNEL_API(lbm_strict)
