// Interface between the two-phase C-ABI layer (lbm2p_api.cu) and its kernels
// (lbm2p_kernels.cu, compiled twice like the single-phase kernels: lbm2p_fast / lbm2p_strict).
#pragma once
#include "lbm_kernels.cuh"

struct Step2Args {
    StepArgs a;               // populations, link words, node classes, rows, flow BCs, force
    // Colour record of a node, read by the colour pass of the next step (the recoloured g_r, g_b of
    // 2phase/lbm_solver_3d_2phase.py:345-363 are re-evaluated from it instead of being stored):
    //   rrb  = (rho_r, rho_b) of the step being closed: OUTPUT of the colour pass and, one step
    //          later, its INPUT (the collision uses exactly these values), hence two buffers;
    //   uq   = (vx, vy, vz, +-q), q = 1 - 1.5 v.v, negative when C != 0 (interface node): written
    //          by the main pass (the collision's velocity);
    //   recC = (Cx, Cy, Cz, 1/|C|), written and read for interface nodes only.
    // 24 B per node (+16 B at the interface) instead of the 152 B of stored g_r, g_b, of which the
    // main pass writes only 16: rho_r, rho_b are never copied into the record.
    float4 *uq;
    const float2 *rrb;        // current (rho_r, rho_b)  [:593-594]
    float2 *rrb_out;          // colour pass: (rho_r, rho_b) of the step it closes
    float4 *recC;
    // node-linear [N] state of the current step (also what to_numpy() shows):
    float *psi;               // :605; solid nodes hold psi_solid so Compute_C needs no solid test
    // dense storage: psi_nb[s] = psi + (ex*ny*nz + ey*nz + ez), the stencil node i + e_s of a node
    // that sits on no lattice face, as "uniform base + node index" (one IMAD.WIDE per load)
    const float *psi_nb[19];
    float psi_solid, CapA;    // :23-24
    float wl, wg, lg0, l1, l2, g1, g2;   // :100-108
    int bc_psi_type[6];       // :34-39
    float bc_psi_val[6];
    // sparse storage (lbm_solver_3d_2phase_sparse.py): compact fluid list + the single-phase pull
    // table (a.flags, a.rb8, a.blk, a.exc, a.lin, a.first, a.count); records, rho_r, rho_b, psi
    // are then indexed by stored node and a solid neighbour of the psi stencil reads psi_solid
    int sparse;
    // dense colour pass: 1 = every node gathers its 19 records (lattices where few nodes are BULK),
    // 0 = records fetched per z-row and handed between lanes (lbm2p_kernels.cu)
    int gather;
};

#define LBM2P_DECLARE_KERNEL_API(NS)                                                          \
    namespace NS {                                                                            \
    cudaError_t launch_main(int mode, const Step2Args &a, int block, cudaStream_t st);        \
    cudaError_t launch_colour(const Step2Args &a, int block, cudaStream_t st);                \
    cudaError_t launch_main_sparse(int mode, const Step2Args &a, cudaStream_t st);            \
    cudaError_t launch_colour_sparse(const Step2Args &a, cudaStream_t st);                    \
    cudaError_t set_inverse_matrix(const float *invM361);                                     \
    }
LBM2P_DECLARE_KERNEL_API(lbm2p_fast)
LBM2P_DECLARE_KERNEL_API(lbm2p_strict)
