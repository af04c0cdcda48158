// Fused D3Q19 MRT kernels for sm_100a.
//
// One launch = one reference step() (Single_phase/LBM_3D_SinglePhase_Solver.py:477-481)
// for a range of nodes, in the order  pull-stream (:259-268, as a pull) -> face BCs
// (:272-370) -> macroscopic (:372-392) -> collision (:222-241) -> store.  The pipeline
// state kept in HBM is the post-collision populations f* (SoA, one plane per direction),
// double buffered (A-B): 19 fp32 read + 19 fp32 written per fluid node = 152 B.
//
//   MODE_STEP     the fused step
//   MODE_EXTRACT  pull + BC + macro only, written to the user-visible rho / v / F arrays
//                 (what to_numpy() shows after the reference's streaming3)
//   MODE_COLLIDE  collision of the user-visible state (first step after init or after a
//                 from_numpy): reads F, rho, v exactly as the reference's colission does
//
// Compiled twice (see lbm_kernels.cuh): namespace lbm_fast / lbm_strict.
#include <cstdlib>

#include "lbm_kernels.cuh"
#include "lbm_sparse_table.cuh"

#ifdef LBM_STRICT
#define LBM_NS lbm_strict
#else
#define LBM_NS lbm_fast
#endif

namespace LBM_NS {
using namespace d3q19;

__device__ __forceinline__ uint32_t vbc_slot(const StepArgs &a, int face, uint32_t lin) {
    const uint32_t z = lin % (uint32_t)a.nz;
    const uint32_t t = lin / (uint32_t)a.nz;
    const uint32_t y = t % (uint32_t)a.ny;
    const uint32_t x = t / (uint32_t)a.ny;
    const uint32_t s = face < 2 ? y * a.nz + z : (face < 4 ? x * a.nz + z : x * a.ny + y);
    return a.vbc_off[face] + s;
}

// Face BCs -> macro -> (collide).  `f` holds the streamed populations on entry.
// force acting on a node (cal_local_force :217-220): the uniform ext_f, or the node's entry of
// the force array (FORCE == 2); `node` indexes the stored order (dense: linear index)
template <int FORCE>
__device__ __forceinline__ void local_force(const StepArgs &a, uint32_t node, float (&frc)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) frc[k] = FORCE >= 2 ? __ldg(a.ff[k] + node) : a.P.force[k];
}
// force of the step being CLOSED by a fused launch (differs only when the array was replaced)
template <int FORCE>
__device__ __forceinline__ void macro_force(const StepArgs &a, uint32_t node, float (&frc)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) frc[k] = FORCE == 3 ? __ldg(a.ffm[k] + node) : (FORCE == 2 ? __ldg(a.ff[k] + node) : a.P.force[k]);
}

// The script copies' fixed-velocity faces (Single_phase/lbm_solver_3d.py:253,268), every one after
// face `after` in order:  F[s] = feq(LR[s],1,u) - F[LR[s]] + feq(s,1,u), IN PLACE for s = 0..18 (a
// later s reads the F[LR[s]] an earlier one has already replaced).
__device__ __forceinline__ void script_velocity_faces(float (&f)[19], const StepArgs &a, uint32_t fl, int after) {
    for (int face = after; face < 6; ++face) {
        if (a.P.bc_type[face] != 2 || !(fl & (FL_AT_X0 << face))) continue;
        const float u0 = a.P.bc_vel[face][0], u1 = a.P.bc_vel[face][1], u2 = a.P.bc_vel[face][2];
#define X(s, ex, ey, ez, o)                                                                    \
    f[s] = feq<o, -(ex), -(ey), -(ez)>(1.0f, u0, u1, u2) - f[o] + feq<s, ex, ey, ez>(1.0f, u0, u1, u2);
        D3Q19_DIRS(X)
#undef X
    }
}

// Boundary_condition :272-370 on the streamed populations of a node that sits on a face with a BC.
// The link word names the face that decides (the last one in the order x0,x1,y0,y1,z0,z1: both forms
// of the class overwrite all 19 populations).  With the script copies' velocity form, which reads
// the current F and so is no overwrite, the word names the last PRESSURE face and every velocity
// face after it is applied in order, found through the at-face bits.  Returns true for a pressure
// node: its velocity of this step goes to vbc[slot] for the next one (:279-281).
__device__ __forceinline__ bool face_bcs(float (&f)[19], const StepArgs &a, uint32_t fl, uint32_t lin, uint32_t &slot) {
    bool pressure = false;
    const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
    int after = 0;
    if (bc) {
        const int face = (int)bc - 1;
        const int type = a.P.bc_type[face];
        if (type == 1) {                      // :274-281  F = feq(rho_bc, v_prev)
            float u0 = 0.f, u1 = 0.f, u2 = 0.f;
            slot = vbc_slot(a, face, lin);
            if (!(fl & FL_PIN_SOLID)) {
                u0 = a.vbc[3 * (size_t)slot + 0];
                u1 = a.vbc[3 * (size_t)slot + 1];
                u2 = a.vbc[3 * (size_t)slot + 2];
            }
            feq_all(f, a.P.bc_rho[face], u0, u1, u2);
            pressure = true;
        } else if (type == 2) {               // :283-288  F = feq(1, bc_vel)
            feq_all(f, 1.0f, a.P.bc_vel[face][0], a.P.bc_vel[face][1], a.P.bc_vel[face][2]);
        }
        after = face + 1;
    }
    if (a.P.vel_bc_script) script_velocity_faces(f, a, fl, after);
    return pressure;
}

template <int FORCE, int MODE>
__device__ __forceinline__ void node_update(float (&f)[19], const StepArgs &a, uint32_t fl,
                                            uint32_t lin, uint32_t node, float &rho, float &ux, float &uy,
                                            float &uz) {
    uint32_t slot = 0;
    bool pressure = false;
    if (a.has_bc) pressure = face_bcs(f, a, fl, lin, slot);
    float frc[3];
    macro_force<FORCE>(a, node, frc);
    macro(f, frc, FORCE != 0, rho, ux, uy, uz);
    if (MODE == MODE_STEP) {
        if (pressure) {
            a.vbc[3 * (size_t)slot + 0] = ux;
            a.vbc[3 * (size_t)slot + 1] = uy;
            a.vbc[3 * (size_t)slot + 2] = uz;
        }
        if (FORCE == 3) local_force<FORCE>(a, node, frc);
        collide(f, a.P, frc, FORCE != 0, rho, ux, uy, uz);
    }
}

// MODE_COLLIDE: the reference's colission() on the user-visible state.
template <int FORCE>
__device__ __forceinline__ void node_collide_only(float (&f)[19], const StepArgs &a, uint32_t fl,
                                                  uint32_t lin, uint32_t node) {
    float rho = 1.0f, ux = 0.f, uy = 0.f, uz = 0.f;
    if (a.F != nullptr) {
#pragma unroll
        for (int s = 0; s < 19; ++s) f[s] = a.F[(size_t)lin * 19 + s];
        rho = a.rho[lin];
        ux = a.v[(size_t)lin * 3 + 0];
        uy = a.v[(size_t)lin * 3 + 1];
        uz = a.v[(size_t)lin * 3 + 2];
    } else {                                      // pristine init(): F = w, rho = 1, v = 0
#pragma unroll
        for (int s = 0; s < 19; ++s) f[s] = weight(s);
    }
    if (a.has_bc) {                               // seed v_prev of pressure-BC nodes
        const uint32_t bc = (fl >> FL_BC_SHIFT) & FL_BC_MASK;
        if (bc && a.P.bc_type[bc - 1] == 1) {
            const uint32_t slot = vbc_slot(a, (int)bc - 1, lin);
            a.vbc[3 * (size_t)slot + 0] = ux;
            a.vbc[3 * (size_t)slot + 1] = uy;
            a.vbc[3 * (size_t)slot + 2] = uz;
        }
    }
    float frc[3];
    local_force<FORCE>(a, node, frc);
    collide(f, a.P, frc, FORCE != 0, rho, ux, uy, uz);
}

template <int MODE>
__device__ __forceinline__ void write_user_fields(const StepArgs &a, uint32_t lin,
                                                  const float (&f)[19], float rho, float ux,
                                                  float uy, float uz) {
    a.rho[lin] = rho;
    a.v[(size_t)lin * 3 + 0] = ux;
    a.v[(size_t)lin * 3 + 1] = uy;
    a.v[(size_t)lin * 3 + 2] = uz;
    if (a.F != nullptr) {
#pragma unroll
        for (int s = 0; s < 19; ++s) a.F[(size_t)lin * 19 + s] = f[s];
    }
}

// ---------------------------------------------------------------------------------------------
// dense lattice: direct addressing, node = linear index i*ny*nz + j*nz + k (z fastest)
// ---------------------------------------------------------------------------------------------
// one node of the dense lattice.  PEER (boundary planes of an x-slab whose neighbours are mapped
// over NVLink): the five populations that cross the cut are also stored straight into the
// neighbour's ghost plane.
template <int FORCE, int MODE, bool SPEC, bool PEER>
__device__ __forceinline__ void dense_node(const StepArgs &a) {
    // blockIdx.x walks the z-chunks of a row (fastest), (blockIdx.z, blockIdx.y) the row groups:
    // blocks that run together then cover whole z-rows, i.e. contiguous 19*nzp*4-byte chunks
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    if (z >= (uint32_t)a.nz || r >= a.row_count) return;
    const uint32_t row = a.row_first + r + (r >= a.row_split ? a.row_skip : 0u);
    const uint32_t idx = row * (uint32_t)a.nz + z;      // node (flags, rho, v, F)
    const uint32_t pidx = row * a.prow + z;             // element inside a population plane
    float f[19];
    bool compute = true;
    float rho = 1.0f, ux = 0.f, uy = 0.f, uz = 0.f;
    if (MODE == MODE_COLLIDE) {
        const uint32_t fl = a.flags[idx];
        if (fl & FL_SOLID) return;
        node_collide_only<FORCE>(f, a, fl, idx, idx);
    } else {
        // Speculative pull (SPEC): the class byte and all 19 populations of the common node
        // (no solid link, no wrap, no BC) are requested together, so a bulk node pays ONE
        // memory latency.  The buffers carry a guard band of a plane + a row at both ends, so
        // the speculative addresses of boundary nodes are always mapped; what they fetch is
        // replaced below.  Without SPEC (mostly-solid lattices) the class is read first.
        const uint8_t cls = a.cls[idx];
        if (!SPEC && cls == NODE_SOLID) return;
        if (SPEC || cls != NODE_SOLID_WRITE) {
#define X(s, ex, ey, ez, o) f[s] = __ldg(a.ppull[s] + pidx);
            D3Q19_DIRS(X)
#undef X
        } else {
#pragma unroll
            for (int s = 0; s < 19; ++s) f[s] = 0.f;
        }
        // solid nodes keep f = F = w, rho = 1, v = 0 (:164-169, :390-392)
        if (SPEC && cls == NODE_SOLID) return;
        compute = cls != NODE_SOLID_WRITE;
        if (MODE == MODE_EXTRACT && !compute) return;
        bool pressure = false;
        uint32_t slot = 0;
        if (cls == NODE_SPECIAL) {
            const uint32_t fl = a.flags[idx];
            // with ghost planes (x-slab) x never wraps; the x bits then mark the GLOBAL faces (BCs)
            const uint32_t flw = a.halo_x ? fl & ~(FL_AT_X0 | FL_AT_X1) : fl;
            const int sx = a.ny * (int)a.prow, sy = (int)a.prow;
            // offsets to the x-1 / x+1 ... neighbours with the periodic wrap of :247-257
            const int oxm = (flw & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
            const int oxp = (flw & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
            const int oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
            const int oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
            const int ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
            const int ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
            // Only the directions the speculation got wrong are fetched again: the pull source
            // is solid -> own opposite population (half-way bounce-back :267-268); the pull
            // crosses a periodic face -> wrapped source (:247-257).
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
#define WRAPS(ex, ey, ez)                                                                      \
    (flw & ((ex > 0 ? FL_AT_X0 : (ex < 0 ? FL_AT_X1 : 0u)) | (ey > 0 ? FL_AT_Y0 : (ey < 0 ? FL_AT_Y1 : 0u)) | \
           (ez > 0 ? FL_AT_Z0 : (ez < 0 ? FL_AT_Z1 : 0u))))
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        if ((fl >> s) & 1u) f[s] = __ldg(a.pown[o] + pidx);                                    \
        else if (WRAPS(ex, ey, ez)) f[s] = __ldg(a.pown[s] + (pidx + OFF(ex, ey, ez)));        \
    }
            D3Q19_DIRS(X)
#undef X
#undef WRAPS
#undef OFF
            if (a.has_bc) pressure = face_bcs(f, a, fl, idx, slot);
        }
        // one copy of the arithmetic for every fluid lane of the warp
        if (compute) {
            float frc[3];
            macro_force<FORCE>(a, idx, frc);
            macro(f, frc, FORCE != 0, rho, ux, uy, uz);
            if (MODE == MODE_STEP) {
                if (pressure) {
                    a.vbc[3 * (size_t)slot + 0] = ux;
                    a.vbc[3 * (size_t)slot + 1] = uy;
                    a.vbc[3 * (size_t)slot + 2] = uz;
                }
                if (FORCE == 3) local_force<FORCE>(a, idx, frc);
                collide(f, a.P, frc, FORCE != 0, rho, ux, uy, uz);
            }
        }
        if (MODE == MODE_EXTRACT) {
            write_user_fields<MODE>(a, idx, f, rho, ux, uy, uz);
            return;
        }
    }
#pragma unroll
    for (int s = 0; s < 19; ++s) a.pout[s][pidx] = f[s];
    if (PEER && compute) {
        // first owned plane (launch rows below row_split): e_x = -1 leaves to the left neighbour's
        // right ghost plane; last owned plane: e_x = +1 to the right neighbour's left ghost plane
        if (r < a.row_split) {
            const long long q = (long long)pidx + a.peer_delta[0];
            a.peer_out[0][0][q] = f[2]; a.peer_out[0][1][q] = f[8]; a.peer_out[0][2][q] = f[10];
            a.peer_out[0][3][q] = f[12]; a.peer_out[0][4][q] = f[14];
        } else {
            const long long q = (long long)pidx + a.peer_delta[1];
            a.peer_out[1][0][q] = f[1]; a.peer_out[1][1][q] = f[7]; a.peer_out[1][2][q] = f[9];
            a.peer_out[1][3][q] = f[11]; a.peer_out[1][4][q] = f[13];
        }
    }
}

template <int FORCE, int MODE, bool SPEC>
__global__ void __launch_bounds__(256) k_dense(const StepArgs a) {
    dense_node<FORCE, MODE, SPEC, false>(a);
}

// ---- grey-scale lattice: partial bounce-back streaming -------------------------------------------
// Grey_Scale/lbm_solver_3d_Macro_Sukop.py:233-247.  The script blends, at every fluid node j and for
// every direction s, the post-collision population with the opposite one of the node ahead,
//     f2[j][s] = f*[j][s] + ns[j] (f*[j + e_s][LR[s]] - f*[j][s]),
// and pushes f2[j][s] to j + e_s -- to every neighbour, there is no bounce-back.  Pulled from node i
// (j = i - e_s, periodic wrap :222-231) that is
//     F[i][s] = f*[j][s] + ns[j] (f*[i][LR[s]] - f*[j][s])        source j fluid
//     F[i][s] = w[s]                                              source j solid (ns >= 1)
// the second line because the script never collides or pushes from a solid node, so that slot of F
// keeps the value init() gave it (tests/test_reference_pin.py::test_grey_scale_script pins this).
// An option off the roofline path: besides the 19 pull sources a node reads its own 19 populations
// and the solid fraction of its 18 neighbours; every node reads its link word, nothing is speculated.
template <int FORCE, int MODE>
__global__ void __launch_bounds__(256) k_dense_grey(const StepArgs a) {
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    if (z >= (uint32_t)a.nz || r >= a.row_count) return;
    const uint32_t row = a.row_first + r;
    const uint32_t idx = row * (uint32_t)a.nz + z;
    const uint32_t pidx = row * a.prow + z;
    const uint32_t fl = a.flags[idx];
    if (fl & FL_SOLID) return;                          // solid nodes keep f = F = w, rho = 1, v = 0
    // steps to the x-1 / x+1 ... neighbours in lattice units, with the periodic wrap
    const int xm = (fl & FL_AT_X0) ? a.nx - 1 : -1, xp = (fl & FL_AT_X1) ? -(a.nx - 1) : 1;
    const int ym = (fl & FL_AT_Y0) ? a.ny - 1 : -1, yp = (fl & FL_AT_Y1) ? -(a.ny - 1) : 1;
    const int zm = (fl & FL_AT_Z0) ? a.nz - 1 : -1, zp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
    const int psy = (int)a.prow, psx = a.ny * psy;      // population planes: rows are prow apart
    const int nsy = a.nz, nsx = a.ny * a.nz;            // node arrays
    float f[19];
#define X(s, ex, ey, ez, o)                                                                    \
    if (s == 0) f[0] = __ldg(a.pown[0] + pidx);                                                \
    else if ((fl >> s) & 1u) f[s] = weight(s);                                                 \
    else {                                                                                     \
        const int dx = ex > 0 ? xm : (ex < 0 ? xp : 0), dy = ey > 0 ? ym : (ey < 0 ? yp : 0),  \
                  dz = ez > 0 ? zm : (ez < 0 ? zp : 0);                                        \
        const float fs = __ldg(a.pown[s] + (pidx + (uint32_t)(dx * psx + dy * psy + dz)));     \
        const float fo = __ldg(a.pown[o] + pidx);                                              \
        const float g = __ldg(a.ns + (idx + (uint32_t)(dx * nsx + dy * nsy + dz)));            \
        f[s] = fs + g * (fo - fs);                                                             \
    }
    D3Q19_DIRS(X)
#undef X
    bool pressure = false;
    uint32_t slot = 0;
    if (a.has_bc) pressure = face_bcs(f, a, fl, idx, slot);
    float rho, ux, uy, uz, frc[3];
    macro_force<FORCE>(a, idx, frc);
    macro(f, frc, FORCE != 0, rho, ux, uy, uz);
    if (MODE == MODE_EXTRACT) {
        write_user_fields<MODE>(a, idx, f, rho, ux, uy, uz);
        return;
    }
    if (pressure) {
        a.vbc[3 * (size_t)slot + 0] = ux;
        a.vbc[3 * (size_t)slot + 1] = uy;
        a.vbc[3 * (size_t)slot + 2] = uz;
    }
    if (FORCE == 3) local_force<FORCE>(a, idx, frc);
    collide(f, a.P, frc, FORCE != 0, rho, ux, uy, uz);
#pragma unroll
    for (int s = 0; s < 19; ++s) a.pout[s][pidx] = f[s];
}

// ---- direct peer-memory halo (x-slabs on GPUs that map each other's buffers) -------------------
// The boundary planes of a slab in ONE kernel that is also the halo exchange: no pack, no
// ncclSend/Recv, no unpack.  Flags in device memory order the ranks (p2p[0], p2p[1]: how many of
// these kernels the left / right neighbour has completed, written BY the neighbour over NVLink;
// p2p[2]: blocks of this launch that are done; p2p[3]: set when a wait timed out):
//   launch k waits until both neighbours have completed k of theirs -- their launch k-1 read the ghost
//   planes this launch overwrites, and wrote the ghost planes this launch reads -- then updates its
//   nodes and stores the crossing populations into the neighbours' ghost planes; the last block to
//   finish publishes k+1 to both neighbours (system-scope release after every block's fence).
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int FORCE, bool SPEC>
__global__ void __launch_bounds__(256) k_dense_peer(const StepArgs a) {
    const bool leader = threadIdx.x == 0 && threadIdx.y == 0;
    if (leader) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.p2p + 0) < a.p2p_launch || ld_acquire_sys(a.p2p + 1) < a.p2p_launch) {
            if (clock64() - t0 > 4000000000ll) {      // ~2 s: a neighbour is gone; do not hang the GPU
                atomicExch(a.p2p + 3, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    dense_node<FORCE, MODE_STEP, SPEC, true>(a);
    __threadfence_system();
    __syncthreads();
    if (leader) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd((unsigned *)a.p2p + 2, 1u) == total - 1u) {
            a.p2p[2] = 0;
            __threadfence_system();
            st_release_sys(a.peer_flag[0], a.p2p_launch + 1);
            st_release_sys(a.peer_flag[1], a.p2p_launch + 1);
        }
    }
}

// Dense lattice stepped IN PLACE on one buffer (AA pattern, lbm_config.sparse = 3): the same two
// alternating steps as the sparse in-place mode (lbm_kernels.cuh).  ODD: pull like k_dense
// (speculative, special nodes re-fetch), then store f*_{LR[s]} back into the very location
// direction s was pulled from -- neighbour slot, own opposite slot (bounce-back) or wrapped slot;
// every location is read and written by exactly one thread.  EVEN: everything a node needs sits
// in its own 19 slots (F_q = slot LR[q]); natural layout again afterwards.  Solid nodes never
// store: a sector written here was read in the same launch, so a partial write merges in L2.
// (The ten directions with e_z != 0 store one element off the 128-byte lines -- two transactions
// per warp store, the odd step takes 1.31 x the even one.  Handing those values to the z neighbour
// through shared memory so that every thread stores at its own z was measured on B200 and is slower,
// 36.2 against 39.1 GLUPS at 256^3: ten STS/LDS pairs, a block barrier and no early exit for solid
// lanes cost more than the split stores.)
template <int FORCE, int MODE, int AA>
__global__ void __launch_bounds__(256, 5) k_dense_aa(const StepArgs a) {
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
    if (z >= (uint32_t)a.nz || r >= a.row_count) return;
    const uint32_t row = a.row_first + r + (r >= a.row_split ? a.row_skip : 0u);
    const uint32_t idx = row * (uint32_t)a.nz + z;
    const uint32_t pidx = row * a.prow + z;
    float f[19];
    const uint8_t cls = a.cls[idx];
    if (AA == AA_EVEN) {
#define X(s, ex, ey, ez, o) f[s] = __ldg(a.pown[o] + pidx);
        D3Q19_DIRS(X)
#undef X
    } else {
#define X(s, ex, ey, ez, o) f[s] = __ldg(a.ppull[s] + pidx);
        D3Q19_DIRS(X)
#undef X
    }
    if (cls == NODE_SOLID || cls == NODE_SOLID_WRITE) return;
    uint32_t fl = 0;
    int oxm = 0, oxp = 0, oym = 0, oyp = 0, ozm = 0, ozp = 0;
    bool pressure = false;
    uint32_t slot = 0;
#define OFF(ex, ey, ez)                                                                        \
    ((ex > 0 ? oxm : (ex < 0 ? oxp : 0)) + (ey > 0 ? oym : (ey < 0 ? oyp : 0)) +               \
     (ez > 0 ? ozm : (ez < 0 ? ozp : 0)))
#define WRAPS(ex, ey, ez)                                                                      \
    (fl & ((ex > 0 ? FL_AT_X0 : (ex < 0 ? FL_AT_X1 : 0u)) | (ey > 0 ? FL_AT_Y0 : (ey < 0 ? FL_AT_Y1 : 0u)) | \
           (ez > 0 ? FL_AT_Z0 : (ez < 0 ? FL_AT_Z1 : 0u))))
    if (cls == NODE_SPECIAL) {
        fl = a.flags[idx];
        if (AA == AA_ODD) {
            const int sx = a.ny * (int)a.prow, sy = (int)a.prow;
            oxm = (fl & FL_AT_X0) ? (a.nx - 1) * sx : -sx;
            oxp = (fl & FL_AT_X1) ? -(a.nx - 1) * sx : sx;
            oym = (fl & FL_AT_Y0) ? (a.ny - 1) * sy : -sy;
            oyp = (fl & FL_AT_Y1) ? -(a.ny - 1) * sy : sy;
            ozm = (fl & FL_AT_Z0) ? (a.nz - 1) : -1;
            ozp = (fl & FL_AT_Z1) ? -(a.nz - 1) : 1;
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        if ((fl >> s) & 1u) f[s] = __ldg(a.pown[o] + pidx);                                    \
        else if (WRAPS(ex, ey, ez)) f[s] = __ldg(a.pown[s] + (pidx + OFF(ex, ey, ez)));        \
    }
            D3Q19_DIRS(X)
#undef X
        }
        if (a.has_bc) pressure = face_bcs(f, a, fl, idx, slot);
    }
    float rho, ux, uy, uz;
    float frc[3];
    macro_force<FORCE>(a, idx, frc);
    macro(f, frc, FORCE != 0, rho, ux, uy, uz);
    if (MODE == MODE_EXTRACT) {
        write_user_fields<MODE>(a, idx, f, rho, ux, uy, uz);
        return;
    }
    if (pressure) {
        a.vbc[3 * (size_t)slot + 0] = ux;
        a.vbc[3 * (size_t)slot + 1] = uy;
        a.vbc[3 * (size_t)slot + 2] = uz;
    }
    if (FORCE == 3) local_force<FORCE>(a, idx, frc);
    collide(f, a.P, frc, FORCE != 0, rho, ux, uy, uz);
    if (AA == AA_EVEN) {
#pragma unroll
        for (int s = 0; s < 19; ++s) a.pout[s][pidx] = f[s];
        return;
    }
    // ODD: back into the pulled locations (pown and pout are the same buffer)
    a.pout[0][pidx] = f[0];
    if (cls != NODE_SPECIAL) {
#define X(s, ex, ey, ez, o) if (s > 0) const_cast<float *>(a.ppull[s])[pidx] = f[o];
        D3Q19_DIRS(X)
#undef X
    } else {
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        if ((fl >> s) & 1u) a.pout[o][pidx] = f[o];                                            \
        else if (WRAPS(ex, ey, ez)) a.pout[s][pidx + OFF(ex, ey, ez)] = f[o];                  \
        else const_cast<float *>(a.ppull[s])[pidx] = f[o];                                     \
    }
        D3Q19_DIRS(X)
#undef X
    }
#undef WRAPS
#undef OFF
}

// ---------------------------------------------------------------------------------------------
// sparse storage: compacted fluid list + 18-neighbour pull table
// ---------------------------------------------------------------------------------------------
#ifndef LBM_SPARSE_MINB
#define LBM_SPARSE_MINB 8
#endif
#ifndef LBM_SPARSE_MINB_ODD
#define LBM_SPARSE_MINB_ODD 6
#endif
// one stored node: pull (through the table slice in shared memory once `bar` flips to
// `parity`), BC, macro, collide, store
template <int FORCE, int MODE, bool COMP, int AA>
__device__ __forceinline__ void sparse_node(const StepArgs &a, SparseTable &s_tab, uint64_t *s_bar,
                                            uint32_t blk, uint32_t i, uint32_t first, uint32_t count,
                                            uint32_t parity = 0u) {
    const bool active = i >= first && i < first + count;
    float f[19];
    uint32_t fl = 0;
    // half-way bounce-back (:267-268): a direction whose pull source is solid takes the node's
    // own opposite population, which sits `stride` elements above or below in the next plane;
    // as ONE signed node index per direction (the context guarantees 2 * stride < 2^31)
    const int32_t ip = (int32_t)i + (int32_t)a.stride, im = (int32_t)i - (int32_t)a.stride;
    if (MODE == MODE_COLLIDE) {
        if (!active) return;
        fl = a.flags[i];
        node_collide_only<FORCE>(f, a, fl, a.lin[i], i);
    } else {
        if (AA == AA_EVEN) {
            // arrival layout: everything this node needs sits in its own 19 slots
            if (!active) return;
            fl = a.has_bc ? a.flags[i] : 0u;
#define X(s, ex, ey, ez, o) f[s] = __ldg(a.pown[o] + i);
            D3Q19_DIRS(X)
#undef X
        } else if (COMP) {
            table_wait(a, blk, s_tab, s_bar, parity);
            if (!active) return;
            fl = s_tab.fl[threadIdx.x];
            if (!(fl & FL_EXCEPTION)) {
                int32_t rb[8];
                table_ranks(s_tab, i, rb);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) f[s] = ldg_gather(a.pown[s] + (((fl >> s) & 1u) ? ((o) > (s) ? ip : im) : comp_source<ex, ey, ez>(i, fl, rb)));
                D3Q19_DIRS(X)
#undef X
            } else {
                const uint32_t slot = table_exc_slot(s_tab);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) f[s] = __ldg(a.pown[s] + (((fl >> s) & 1u) ? ((o) > (s) ? ip : im) : __ldg(a.exc[s > 0 ? s - 1 : 0] + slot)));
                D3Q19_DIRS(X)
#undef X
            }
            // the rest population last: requested first, ptxas parks it in local memory at once,
            // which waits for it to arrive before any gather is issued
            f[0] = __ldg(a.pown[0] + i);
        } else {
            if (!active) return;
            fl = a.has_bc ? a.flags[i] : 0u;
            f[0] = __ldg(a.pown[0] + i);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        const int32_t j = __ldg(a.nbr[s > 0 ? s - 1 : 0] + i);                                 \
        f[s] = j >= 0 ? __ldg(a.pown[s] + (uint32_t)j) : __ldg(a.pown[o] + i);                 \
    }
            D3Q19_DIRS(X)
#undef X
        }
        float rho, ux, uy, uz;
        const uint32_t lin = (MODE == MODE_EXTRACT || a.has_bc) ? a.lin[i] : 0u;
        node_update<FORCE, MODE>(f, a, fl, lin, i, rho, ux, uy, uz);
        if (MODE == MODE_EXTRACT) {
            write_user_fields<MODE>(a, lin, f, rho, ux, uy, uz);
            return;
        }
    }
    if (AA == AA_ODD && MODE == MODE_STEP && COMP) {
        // in place: f*_{LR[s]}(i) goes back to the location direction s was pulled from (each
        // location is read and written by exactly this one thread, loads above, stores here)
        a.pout[0][i] = f[0];
        if (!(fl & FL_EXCEPTION)) {
            int32_t rb[8];
            table_ranks(s_tab, i, rb);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) a.pout[s][((fl >> s) & 1u) ? ((o) > (s) ? ip : im) : comp_source<ex, ey, ez>(i, fl, rb)] = f[o];
            D3Q19_DIRS(X)
#undef X
        } else {
            const uint32_t slot = table_exc_slot(s_tab);
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) a.pout[s][((fl >> s) & 1u) ? ((o) > (s) ? ip : im) : __ldg(a.exc[s > 0 ? s - 1 : 0] + slot)] = f[o];
            D3Q19_DIRS(X)
#undef X
        }
    } else {
#pragma unroll
        for (int s = 0; s < 19; ++s) a.pout[s][i] = f[s];
    }
}

// One 256-node table block per CTA.  (Measured again in round 2 and dropped again: several blocks
// per CTA with the next slices prefetched into a 3-slot ring, full / empty mbarriers instead of a
// block barrier -- 0.93 ms per step against 0.73 on the 512^3 pack.  Independent CTAs hide the
// table fetch better than any coupling of the warps of one CTA.)
// occupancy: 8 blocks (32 registers) per SM, except the in-place odd step, which keeps the 19
// pull locations live across the collision and is faster unspilled at 6 blocks (40 registers)
template <int FORCE, int MODE, bool COMP, int AA>
__global__ void __launch_bounds__(SPARSE_BLOCK, (AA == AA_ODD && MODE == MODE_STEP) ? LBM_SPARSE_MINB_ODD : LBM_SPARSE_MINB)
k_sparse(const StepArgs a) {
    constexpr bool TABLE = COMP && MODE != MODE_COLLIDE && AA != AA_EVEN;   // phase 1 needed
    __shared__ SparseTable s_tab;
    __shared__ uint64_t s_bar;
    uint32_t first = a.first, count = a.count, b = blockIdx.x;
    if (b >= a.nb1) {                // second range of a two-range launch (StepArgs::nb1)
        b -= a.nb1;
        first = a.first2;
        count = a.count2;
    }
    const uint32_t blk = b + first / SPARSE_BLOCK;
    if (TABLE) {
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            table_fetch(a, blk, s_tab, &s_bar);
            if (MODE == MODE_STEP && a.prefetch_dist) {
                // ask L2 for the table slice of the block one wave ahead
                const uint32_t pb = blk * SPARSE_BLOCK + a.prefetch_dist;
                if (pb + SPARSE_BLOCK <= first + count) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.rb8[k] + pb), "r"(SPARSE_BLOCK * 1u) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.flags + pb), "r"(SPARSE_BLOCK * 4u) : "memory");
                }
            }
        }
    }
    sparse_node<FORCE, MODE, COMP, AA>(a, s_tab, &s_bar, blk, blk * SPARSE_BLOCK + threadIdx.x, first, count);
}

template <int FORCE, int MODE>
static void launch_dense_t(const StepArgs &a, int block, cudaStream_t st) {
    // block = (BX along z, BY rows); BX = nz rounded up to a warp, capped at `block`
    // split a z-row into the fewest chunks of at most `block` threads, of equal (warp-rounded) size
    const int nchunk = (a.nz + block - 1) / block;
    int bx = ((a.nz + nchunk - 1) / nchunk + 31) / 32 * 32;
    if (bx > block) bx = block;
    int by = block / bx;
    if (by < 1) by = 1;
    dim3 blk(bx, by, 1);
    const unsigned rg = (a.row_count + by - 1) / by;
    const unsigned gy = rg < 32768u ? rg : 32768u;
    dim3 grid((a.nz + bx - 1) / bx, gy, (rg + gy - 1) / gy);
    if (a.ns != nullptr && MODE != MODE_COLLIDE) {                  // grey-scale lattice (two buffers, no slabs)
        k_dense_grey<FORCE, MODE == MODE_COLLIDE ? MODE_STEP : MODE><<<grid, blk, 0, st>>>(a);
        return;
    }
    if (a.aa != AA_OFF && MODE != MODE_COLLIDE) {
        if (a.aa == AA_ODD)
            k_dense_aa<FORCE, MODE == MODE_COLLIDE ? MODE_STEP : MODE, AA_ODD><<<grid, blk, 0, st>>>(a);
        else
            k_dense_aa<FORCE, MODE == MODE_COLLIDE ? MODE_STEP : MODE, AA_EVEN><<<grid, blk, 0, st>>>(a);
        return;
    }
    if (MODE == MODE_STEP && FORCE < 2 && a.p2p != nullptr) {      // boundary planes + peer-memory halo
        if (a.spec)
            k_dense_peer<FORCE < 2 ? FORCE : 0, true><<<grid, blk, 0, st>>>(a);
        else
            k_dense_peer<FORCE < 2 ? FORCE : 0, false><<<grid, blk, 0, st>>>(a);
        return;
    }
    if (a.spec)
        k_dense<FORCE, MODE, true><<<grid, blk, 0, st>>>(a);
    else
        k_dense<FORCE, MODE, false><<<grid, blk, 0, st>>>(a);
}

template <int FORCE, int MODE>
static void launch_sparse_t(const StepArgs &a, int block, cudaStream_t st) {
    (void)block;
    block = SPARSE_BLOCK;
    const unsigned b0 = a.first / SPARSE_BLOCK, b1 = (a.first + a.count + SPARSE_BLOCK - 1) / SPARSE_BLOCK;
    unsigned grid = b1 - b0;
    if (a.nb1 != 0xFFFFFFFFu)        // two ranges: nb1 blocks of the first, then the second
        grid = a.nb1 + ((a.first2 + a.count2 + SPARSE_BLOCK - 1) / SPARSE_BLOCK - a.first2 / SPARSE_BLOCK);
    if (!a.compressed)
        k_sparse<FORCE, MODE, false, AA_OFF><<<grid, block, 0, st>>>(a);
    else if (a.aa == AA_ODD)
        k_sparse<FORCE, MODE, true, AA_ODD><<<grid, block, 0, st>>>(a);
    else if (a.aa == AA_EVEN)
        k_sparse<FORCE, MODE, true, AA_EVEN><<<grid, block, 0, st>>>(a);
    else
        k_sparse<FORCE, MODE, true, AA_OFF><<<grid, block, 0, st>>>(a);
}

#define DISPATCH(FN)                                                                           \
    switch (a.force * 4 + mode) {                                                              \
        case 0: FN<0, MODE_STEP>(a, block, st); break;                                         \
        case 1: FN<0, MODE_EXTRACT>(a, block, st); break;                                      \
        case 2: FN<0, MODE_COLLIDE>(a, block, st); break;                                      \
        case 4: FN<1, MODE_STEP>(a, block, st); break;                                         \
        case 5: FN<1, MODE_EXTRACT>(a, block, st); break;                                      \
        case 6: FN<1, MODE_COLLIDE>(a, block, st); break;                                      \
        case 8: FN<2, MODE_STEP>(a, block, st); break;                                         \
        case 9: FN<2, MODE_EXTRACT>(a, block, st); break;                                      \
        case 10: FN<2, MODE_COLLIDE>(a, block, st); break;                                     \
        case 12: FN<3, MODE_STEP>(a, block, st); break;                                        \
        default: return cudaErrorInvalidValue;                                                 \
    }

static cudaError_t launch_any(bool sparse, int mode, const StepArgs &a, int block, cudaStream_t st) {
    if (block <= 0 || block > 256 || block % 32) block = 256;
    if (sparse) {
        if (a.count == 0 && (a.nb1 == 0xFFFFFFFFu || a.count2 == 0)) return cudaSuccess;
        DISPATCH(launch_sparse_t)
    } else {
        if (a.row_count == 0) return cudaSuccess;
        DISPATCH(launch_dense_t)
    }
    return cudaGetLastError();
}

cudaError_t launch_dense(int mode, const StepArgs &a, int block, cudaStream_t st) {
    return launch_any(false, mode, a, block, st);
}
cudaError_t launch_sparse(int mode, const StepArgs &a, int block, cudaStream_t st) {
    return launch_any(true, mode, a, block, st);
}

cudaError_t set_inverse_matrix(const float *invM361) {
#ifdef LBM_STRICT
    return cudaMemcpyToSymbol(c_invM, invM361, 361 * sizeof(float));
#else
    (void)invM361;
    return cudaSuccess;
#endif
}

}  // namespace LBM_NS
