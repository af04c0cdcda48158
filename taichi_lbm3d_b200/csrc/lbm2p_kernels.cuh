// Interface between the two-phase C-ABI layer (lbm2p_api.cu) and its kernels
// (lbm2p_kernels.cu, compiled twice like the single-phase kernels: lbm2p_fast / lbm2p_strict).
#pragma once
#include "lbm_kernels.cuh"

enum { COLOUR_SPEC_NONE = 0 };

struct Step2Args {
    StepArgs a;               // populations, link words, node classes, rows, flow BCs, force
    // colour record published by the collision of step n, read by the colour pass:
    //   0 rho_r, 1 rho_b, 2..4 v, 5..7 C   -- the values the collision used (2phase/
    //   lbm_solver_3d_2phase.py:345-363), node-linear [8][N] with a guard band
    float *rec[8];
    // node-linear [N] state of the current step (also what to_numpy() shows):
    float *rho_r, *rho_b;     // :593-594
    float *psi;               // :605; solid nodes hold psi_solid so Compute_C needs no solid test
    float psi_solid, CapA;    // :23-24
    float wl, wg, lg0, l1, l2, g1, g2;   // :100-108
    int bc_psi_type[6];       // :34-39
    float bc_psi_val[6];
};

#define LBM2P_DECLARE_KERNEL_API(NS)                                                          \
    namespace NS {                                                                            \
    cudaError_t launch_main(int mode, const Step2Args &a, int block, cudaStream_t st);        \
    cudaError_t launch_colour(const Step2Args &a, int block, cudaStream_t st);                \
    cudaError_t set_inverse_matrix(const float *invM361);                                     \
    }
LBM2P_DECLARE_KERNEL_API(lbm2p_fast)
LBM2P_DECLARE_KERNEL_API(lbm2p_strict)
