"""Multi-GPU x-slab decomposition of the single-phase solver (one process per GPU).

The reference is single-device (no collective or decomposition code anywhere in its tree);
this is the new work BASELINE.json asks for.  The domain is cut into contiguous x-slabs, one
per rank, each with one ghost plane on either side.  Streaming (reference
Single_phase/LBM_3D_SinglePhase_Solver.py:259-268) couples a node only to its 18 neighbours,
so per step and per cut exactly the five populations with e_x = +1 (s = 1,7,9,11,13) travel
right and the five with e_x = -1 (s = 2,8,10,12,14) travel left: 20 bytes per face node per
direction.  Collision, boundary conditions and macroscopic sums are node-local.

Layering
  SlabPartition   pure host arithmetic (who owns which planes, who the neighbours are)
  HaloExchanger   the per-step exchange over torch.distributed point-to-point ops
                  (NCCL on GPUs; the CPU tests drive the same code over gloo)
  SlabSolver      the user-facing object: reference-style setters, step/run, gathers

Schedule of one step (overlap):  boundary planes first -> their outgoing populations are
packed and exchanged on a side stream while the interior planes are updated -> join.
"""
import ctypes

import numpy as np

from . import _lib
from .LBM_3D_SinglePhase_Solver import LB3D_Solver_Single_Phase


class SlabPartition:
    """Contiguous x-slabs over `world` ranks; the ring is periodic like periodic_index (:250-251)."""

    def __init__(self, gnx, world, rank):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad world/rank")
        if gnx < world:
            raise ValueError("fewer x planes (%d) than ranks (%d)" % (gnx, world))
        self.gnx, self.world, self.rank = gnx, world, rank
        base, rem = divmod(gnx, world)
        sizes = [base + (1 if r < rem else 0) for r in range(world)]
        self.sizes = sizes
        self.x0 = sum(sizes[:rank])
        self.x1 = self.x0 + sizes[rank]           # owned planes [x0, x1)
        self.left = (rank - 1) % world
        self.right = (rank + 1) % world
        self.own = sizes[rank]
        self.local_nx = self.own + 2               # + ghost plane on each side

    @property
    def x_face_mask(self):
        """bit0: this slab holds the global x0 face; bit1: it holds the global x1 face"""
        return (1 if self.rank == 0 else 0) | (2 if self.rank == self.world - 1 else 0)

    def local_planes(self):
        """global x index of every local plane, ghosts included (periodic wrap)"""
        return [(x % self.gnx) for x in range(self.x0 - 1, self.x1 + 1)]

    def local_solid(self, global_solid):
        g = np.asarray(global_solid)
        if g.shape[0] != self.gnx:
            raise ValueError("global geometry has %d planes, expected %d" % (g.shape[0], self.gnx))
        return np.ascontiguousarray(np.take(g, self.local_planes(), axis=0))

    def owned(self, local_array):
        """strip the ghost planes of a local field"""
        return local_array[1:-1]


class HaloExchanger:
    """Exchanges the face-crossing populations of the current post-collision state.

    `backend` must provide
        pack(side) -> tensor          side 0: e_x=-1 populations of the first owned plane
                                      side 1: e_x=+1 populations of the last owned plane
        unpack(side, tensor)          side 0: left ghost plane <- what the left rank packed(1)
                                      side 1: right ghost plane <- what the right rank packed(0)
        recv_buffer(side) -> tensor   staging buffer sized for that ghost plane
    `dist` is torch.distributed (or None for a single rank: the ring closes on itself).
    """

    def __init__(self, part, backend, dist=None):
        self.part, self.backend, self.dist = part, backend, dist

    def exchange(self, which=0):
        p, b = self.part, self.backend
        to_left = b.pack(0, which)
        to_right = b.pack(1, which)
        if p.world == 1 or self.dist is None:
            # periodic ring of one slab: my right face feeds my own left ghost and v.v.
            b.unpack(0, to_right, which)
            b.unpack(1, to_left, which)
            return
        dist = self.dist
        from_left = b.recv_buffer(0)
        from_right = b.recv_buffer(1)
        # Order matters when left == right (two ranks): each pair of ranks matches sends to
        # receives in posting order, and the peer posts its "to_right" first as well.
        ops = [dist.P2POp(dist.isend, to_right, p.right), dist.P2POp(dist.isend, to_left, p.left),
               dist.P2POp(dist.irecv, from_left, p.left), dist.P2POp(dist.irecv, from_right, p.right)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        b.unpack(0, from_left, which)
        b.unpack(1, from_right, which)


class _LocalSlab(LB3D_Solver_Single_Phase):
    """The slab of one rank: a halo_x context of local_nx planes."""

    def __init__(self, part, ny, nz, **kw):
        super().__init__(part.local_nx, ny, nz, **kw)
        self._part = part

    def _config(self):
        cfg = super()._config()
        cfg.halo_x = 1
        cfg.x_face_mask = self._part.x_face_mask
        return cfg


class _CudaBackend:
    """pack/unpack through the C ABI (lbm_halo_pack / lbm_halo_unpack) with torch staging buffers."""

    def __init__(self, slab):
        import torch
        self.torch = torch
        self.slab = slab
        lib, ctx = slab._lib, slab._ctx
        self.counts = [int(lib.lbm_halo_count(ctx, pl)) for pl in range(4)]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.send = [torch.empty(5 * max(self.counts[1], 1), dtype=torch.float32, device=dev),
                     torch.empty(5 * max(self.counts[2], 1), dtype=torch.float32, device=dev)]
        self.recv = [torch.empty(5 * max(self.counts[0], 1), dtype=torch.float32, device=dev),
                     torch.empty(5 * max(self.counts[3], 1), dtype=torch.float32, device=dev)]

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def pack(self, side, which=0):
        s = self.slab
        s._ck(s._lib.lbm_halo_pack(s._ctx, side, which, ctypes.c_void_p(self.send[side].data_ptr()), self._stream()),
              "lbm_halo_pack")
        return self.send[side]

    def unpack(self, side, tensor, which=0):
        s = self.slab
        s._ck(s._lib.lbm_halo_unpack(s._ctx, side, which, ctypes.c_void_p(tensor.data_ptr()), self._stream()),
              "lbm_halo_unpack")

    def recv_buffer(self, side):
        return self.recv[side]


class SlabSolver:
    """D3Q19 MRT single-phase solver over `world_size` GPUs (x-slabs); same setters as
    LB3D_Solver_Single_Phase.  Uses the default process group of torch.distributed when it
    is initialised, otherwise runs as a single slab whose ring closes on itself."""

    def __init__(self, nx, ny, nz, sparse_storage=False, strict=False, tau_mode="class", overlap=True,
                 transport="native", guo_mode="class", vel_bc_mode="class"):
        """transport="native": the step loop and the ncclSend/ncclRecv halo exchange run inside
        the C library (lbm_run_slab), a handful of launches per step and no Python in between;
        transport="torch": the same schedule driven from Python over torch.distributed P2P ops
        (the path the gloo CPU tests exercise)."""
        if transport not in ("native", "torch"):
            raise ValueError("transport must be 'native' or 'torch'")
        self.transport = transport
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        world = self.dist.get_world_size() if self.dist else 1
        rank = self.dist.get_rank() if self.dist else 0
        self.nx, self.ny, self.nz = nx, ny, nz
        self.part = SlabPartition(nx, world, rank)
        self.local = _LocalSlab(self.part, ny, nz, sparse_storage=sparse_storage, strict=strict,
                                tau_mode=tau_mode, guo_mode=guo_mode, vel_bc_mode=vel_bc_mode)
        self.overlap = bool(overlap) and self.part.own >= 3
        self._started = False
        self._comm_stream = None
        self.peer_memory = False
        # reference-style setters are forwarded to the local slab (x faces only act on the
        # ranks that hold them, through x_face_mask)
        for name in dir(LB3D_Solver_Single_Phase):
            if name.startswith("set_bc_") or name in ("set_force", "set_viscosity"):
                setattr(self, name, getattr(self.local, name))

    # ---- setup -------------------------------------------------------------------------------
    def set_solid(self, global_solid):
        self._global_solid = (np.asarray(global_solid) > 0).view(np.int8)
        self.local.solid.from_numpy(self.part.local_solid(self._global_solid))

    def set_local_solid(self, local_solid_with_ghosts):
        """for domains too large to hold on every host: planes [x0-1 .. x1] of this rank"""
        self.local.solid.from_numpy(local_solid_with_ghosts)

    def set_force_field(self, global_force):
        """per-node force (nx, ny, nz, 3) of the whole domain, or None: every rank keeps its planes"""
        if global_force is None:
            self.local.set_force_field(None)
            return
        g = np.asarray(global_force, np.float32)
        self.local.set_force_field(np.ascontiguousarray(np.take(g, self.part.local_planes(), axis=0)))

    def init_simulation(self):
        self.local.init_simulation()
        self.backend = _CudaBackend(self.local)
        self.halo = HaloExchanger(self.part, self.backend, self.dist)
        self._comm_stream = self.torch.cuda.Stream(priority=-1)     # ahead of the interior kernel
        self._started = False
        if self.transport == "native":
            loc = self.local
            ident = (ctypes.c_char * 128)()
            if self.part.world > 1:
                if self.dist.get_backend() != "nccl":
                    raise _lib.LbmError("transport='native' needs the nccl process group")
                # one NCCL communicator per process (lbm_nccl.cuh): only the first solver needs an id
                if not loc._lib.lbm_comm_ready(self.part.world, self.part.rank):
                    box = [None]
                    if self.part.rank == 0:
                        loc._ck(loc._lib.lbm_comm_unique_id(ident), "lbm_comm_unique_id")
                        box[0] = bytes(ident.raw)
                    self.dist.broadcast_object_list(box, src=0)
                    ident.raw = box[0]
            loc._ck(loc._lib.lbm_comm_init(loc._ctx, ident, self.part.world, self.part.rank), "lbm_comm_init")
            # Mapping the neighbours' buffers costs 50 ms (2 GPUs) to 200 ms (8): it is done at the first
            # multi-step run(n), which announces a job long enough to repay it; a job of a few single
            # step() calls stays on the NCCL exchange.  LBM3D_P2P=1 maps here, LBM3D_P2P=0 never.
            import os
            self._peer_tried = False
            if os.environ.get("LBM3D_P2P", "") == "1":
                self.peer_memory, self._peer_tried = self._connect_peer_memory(), True

    def _connect_peer_memory(self):
        """Direct peer-memory halo (include/lbm3d.h: lbm_p2p_*): every rank maps its neighbours'
        population buffers (CUDA IPC over NVLink) and the boundary-plane kernel writes the crossing
        populations straight into their ghost planes.  Used when EVERY rank could connect (dense
        storage, overlapped schedule, one node); LBM3D_P2P=0 keeps the NCCL exchange."""
        import os
        loc, dist, part = self.local, self.dist, self.part
        if part.world < 2 or not self.overlap or loc.sparse_storage or os.environ.get("LBM3D_P2P", "") == "0":
            return False
        lib, ctx = loc._lib, loc._ctx
        torch = self.torch
        blob = (ctypes.c_char * 256)()
        exported = lib.lbm_p2p_export(ctx, blob) == 0
        # one all_gather of 257 bytes per rank: the blob and whether it is valid
        mine = torch.frombuffer(bytearray(bytes(blob.raw) + (b"\x01" if exported else b"\x00")), dtype=torch.uint8).cuda()
        everyone = torch.empty(part.world * 257, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(everyone, mine)
        everyone = everyone.cpu().numpy().reshape(part.world, 257)
        ok = bool(everyone[:, 256].all())
        if ok:
            left = (ctypes.c_char * 256).from_buffer_copy(everyone[part.left, :256].tobytes())
            right = (ctypes.c_char * 256).from_buffer_copy(everyone[part.right, :256].tobytes())
            ok = lib.lbm_p2p_connect(ctx, left, right) == 0
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks or none: the two exchanges do not mix
        ok = bool(int(flag.item()))
        if ok:
            loc._ck(lib.lbm_p2p_enable(ctx, 1), "lbm_p2p_enable")
        return ok

    def close(self):
        """collective: unmap the neighbours' buffers on every rank, then release the slab.  A solver
        that used the peer-memory halo must be closed (not just dropped) before another is built:
        memory exported to a neighbour has to outlive the neighbour's mapping of it."""
        loc = self.local
        if loc._ctx is not None and self.peer_memory:
            loc._lib.lbm_p2p_disconnect(loc._ctx)
            self.peer_memory = False
            if self.dist:
                self.dist.barrier()
        if loc._ctx is not None:
            loc._lib.lbm_destroy(loc._ctx)
            loc._ctx = None

    # ---- stepping ------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _begin(self):
        """first collision of the user-visible state + first halo exchange"""
        loc = self.local
        r = loc._lib.lbm_step_begin(loc._ctx, self._stream())
        loc._ck(r, "lbm_step_begin")
        self.halo.exchange(0)
        self._started = True

    def _one_step(self):
        loc, torch = self.local, self.torch
        lib, ctx = loc._lib, loc._ctx
        own = self.part.own
        if not self.overlap:
            loc._ck(lib.lbm_step(ctx, 1, self._stream()), "lbm_step")
            self.halo.exchange(0)
            return
        main = torch.cuda.current_stream()
        # 1. boundary planes (local x = 1 and x = own) into the NEXT buffer
        loc._ck(lib.lbm_step_planes(ctx, 1, 2, self._stream()), "lbm_step_planes")
        loc._ck(lib.lbm_step_planes(ctx, own, own + 1, self._stream()), "lbm_step_planes")
        ev = torch.cuda.Event()
        ev.record(main)
        # 2. exchange their outgoing populations on the side stream ...
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            self.halo.exchange(1)
            done = torch.cuda.Event()
            done.record(self._comm_stream)
        # 3. ... while the interior planes are updated
        loc._ck(lib.lbm_step_planes(ctx, 2, own, self._stream()), "lbm_step_planes")
        main.wait_event(done)
        loc._ck(lib.lbm_step_flip(ctx), "lbm_step_flip")

    def step(self):
        self.run(1)

    def run(self, nsteps):
        n = int(nsteps)
        if n <= 0:
            return
        if self.transport == "native":
            loc = self.local
            if n > 1 and not getattr(self, "_peer_tried", True):
                self.peer_memory, self._peer_tried = self._connect_peer_memory(), True
            loc._ck(loc._lib.lbm_run_slab(loc._ctx, n, int(self.overlap), self._stream()), "lbm_run_slab")
            self._started = True
            return
        if not self._started:
            self._begin()
            n -= 1
        for _ in range(n):
            self._one_step()

    def invalidate(self):
        """call after assigning fields of the local slab (from_numpy): the pipeline restarts"""
        self._started = False

    # ---- results ---------------------------------------------------------------------------------
    @property
    def launch_count(self):
        return self.local.launch_count

    def get_max_v(self):
        """max |v| over the whole domain (cal_max_v :399-402): device reduction per slab, then
        MAX over the ranks.  Ghost planes are never written by a step and hold v = 0 from
        init_simulation, so the reduction may run over the whole local lattice."""
        m = self.local.get_max_v()
        if self.dist:
            t = self.torch.tensor([m], dtype=self.torch.float32,
                                  device="cuda" if self.dist.get_backend() == "nccl" else "cpu")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            m = float(t.item())
        return m

    def local_field(self, name, out=None):
        """owned planes of rho / v / F / solid on this rank, shape (own, ny, nz[, C]).  `out`: a
        C-contiguous float32 buffer for the WHOLE local slab, ghost planes included, shape
        (own + 2, ny, nz[, C]) -- e.g. pinned host memory -- whose owned part is returned as a view."""
        return self.part.owned(getattr(self.local, name).to_numpy(out=out) if out is not None
                               else getattr(self.local, name).to_numpy())

    def gather_field(self, name, dst=0):
        """the global field on rank `dst` (None elsewhere)"""
        return _gather_planes(self, self.local_field(name), dst)


def _gather_planes(solver, loc, dst):
    """concatenate the owned planes of every rank on rank `dst`: one tensor gather over the process
    group (slabs padded to the largest), no pickling of whole fields"""
    dist, part = solver.dist, solver.part
    if not dist:
        return loc
    torch = solver.torch
    on_gpu = dist.get_backend() == "nccl"
    most = max(part.sizes)
    t = torch.zeros((most,) + loc.shape[1:], dtype=torch.from_numpy(loc[:0]).dtype)
    t[:loc.shape[0]] = torch.from_numpy(np.ascontiguousarray(loc))
    if on_gpu:
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(part.world)] if part.rank == dst else None
    dist.gather(t, parts, dst=dst)
    if part.rank != dst:
        return None
    return np.concatenate([p.cpu().numpy()[:n] for p, n in zip(parts, part.sizes)], axis=0)


# ---------------------------------------------------------------------------------------------
# two-phase colour-gradient solver over x-slabs
# ---------------------------------------------------------------------------------------------
def _two_phase_slab_class():
    from .lbm_solver_3d_2phase import LB3D_Solver_Two_Phase

    class _Slab(LB3D_Solver_Two_Phase):
        def __init__(self, part, ny, nz, **kw):
            super().__init__(part.local_nx, ny, nz, **kw)
            self._part = part

        def _config_flags(self):
            m = self._part.x_face_mask
            return 1 | (2 if m & 1 else 0) | (4 if m & 2 else 0) | (8 if self.sparse_storage else 0)
            # LBM2P_HALO_X | HOLDS_X0 | HOLDS_X1 | LBM2P_SPARSE

    return _Slab


class StagedHaloExchanger:
    """The two exchanges of a two-phase step over torch.distributed point-to-point ops.

    `backend` provides, for stage 0 (populations + colour records) and stage 1 (psi):
        pack(stage, side) -> tensor        side 0: the first owned plane (goes to the left rank)
                                           side 1: the last owned plane (goes to the right rank)
        unpack(stage, side, tensor)        side 0: left ghost plane <- what the left rank packed(1)
                                           side 1: right ghost plane <- what the right rank packed(0)
        recv_buffer(stage, side) -> tensor
    `dist` is torch.distributed (None or world 1: the ring closes on itself)."""

    def __init__(self, part, backend, dist=None):
        self.part, self.backend, self.dist = part, backend, dist

    def exchange(self, stage):
        p, b = self.part, self.backend
        to_left = b.pack(stage, 0)
        to_right = b.pack(stage, 1)
        if p.world == 1 or self.dist is None:
            b.unpack(stage, 0, to_right)
            b.unpack(stage, 1, to_left)
            return
        dist = self.dist
        from_left, from_right = b.recv_buffer(stage, 0), b.recv_buffer(stage, 1)
        # posting order matters when left == right (two ranks): pairs match in order
        ops = [dist.P2POp(dist.isend, to_right, p.right), dist.P2POp(dist.isend, to_left, p.left),
               dist.P2POp(dist.irecv, from_left, p.left), dist.P2POp(dist.irecv, from_right, p.right)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        b.unpack(stage, 0, from_left)
        b.unpack(stage, 1, from_right)


class _Cuda2PBackend:
    """pack / unpack through the C ABI (lbm2p_halo_pack / lbm2p_halo_unpack), torch staging buffers"""

    def __init__(self, slab, torch):
        self.slab, self.torch = slab, torch
        lib, ctx = slab._lib, slab._ctx
        dev = torch.device("cuda", torch.cuda.current_device())
        self.count = [int(lib.lbm2p_halo_floats(ctx, 0)), int(lib.lbm2p_halo_floats(ctx, 1))]
        # nodes of the four halo planes (a sparse slab's planes hold their fluid nodes only)
        self.nodes = [int(lib.lbm2p_halo_count(ctx, pl)) for pl in range(4)]
        self.per_node = (13, 3)
        self.send = [torch.empty(self.count[0] + 4, dtype=torch.float32, device=dev) for _ in range(2)]
        self.recv = [torch.empty(self.count[0] + 4, dtype=torch.float32, device=dev) for _ in range(2)]

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def pack(self, stage, side):
        s = self.slab
        s._ck(s._lib.lbm2p_halo_pack(s._ctx, stage, side, ctypes.c_void_p(self.send[side].data_ptr()), self._stream()),
              "lbm2p_halo_pack")
        return self.send[side][:self.per_node[stage] * self.nodes[1 if side == 0 else 2]]

    def unpack(self, stage, side, tensor):
        s = self.slab
        s._ck(s._lib.lbm2p_halo_unpack(s._ctx, stage, side, ctypes.c_void_p(tensor.data_ptr()), self._stream()),
              "lbm2p_halo_unpack")

    def recv_buffer(self, stage, side):
        return self.recv[side][:self.per_node[stage] * self.nodes[0 if side == 0 else 3]]


class TwoPhaseSlabSolver:
    """Two-phase colour-gradient solver over `world_size` GPUs (x-slabs).  The attributes of
    LB3D_Solver_Two_Phase (niu_l, CapA, bc_psi_x_left, ...) are set on ``.local``; geometry and
    phase field are given globally (set_fields) or per slab (set_local_fields).

    Per step and cut (include/lbm3d_2phase.h): after the colour pass psi of the boundary plane
    (4 B per face node), after the main pass the 5 crossing populations and the 24(+16)-byte
    colour record (60 B per face node).  transport="native": ncclSend/ncclRecv inside the C
    library (lbm2p_run_slab); "torch": the same schedule over torch.distributed P2P ops."""

    def __init__(self, nx, ny, nz, strict=False, transport="native", sparse_storage=False):
        if transport not in ("native", "torch"):
            raise ValueError("transport must be 'native' or 'torch'")
        import torch
        import torch.distributed as dist
        self.torch, self.transport = torch, transport
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        world = self.dist.get_world_size() if self.dist else 1
        rank = self.dist.get_rank() if self.dist else 0
        self.nx, self.ny, self.nz = nx, ny, nz
        self.part = SlabPartition(nx, world, rank)
        self.local = _two_phase_slab_class()(self.part, ny, nz, strict=strict, sparse_storage=sparse_storage)
        self._started = False

    # ---- setup -------------------------------------------------------------------------------
    def set_fields(self, global_solid, global_psi):
        self.local.solid.from_numpy(self.part.local_solid(global_solid))
        psi = np.asarray(global_psi, np.float32)
        self.local.psi.from_numpy(np.ascontiguousarray(np.take(psi, self.part.local_planes(), axis=0)))

    def set_local_fields(self, local_solid_with_ghosts, local_psi_with_ghosts):
        self.local.solid.from_numpy(local_solid_with_ghosts)
        self.local.psi.from_numpy(local_psi_with_ghosts)

    def init_simulation(self):
        loc = self.local
        loc.init_simulation()
        lib, ctx = loc._lib, loc._ctx
        self._started = False
        self.backend = _Cuda2PBackend(loc, self.torch)
        self.halo = StagedHaloExchanger(self.part, self.backend, self.dist)
        if self.transport == "native":
            ident = (ctypes.c_char * 128)()
            if self.part.world > 1:
                if self.dist.get_backend() != "nccl":
                    raise _lib.LbmError("transport='native' needs the nccl process group")
                if not lib.lbm_comm_ready(self.part.world, self.part.rank):      # one communicator per process
                    box = [None]
                    if self.part.rank == 0:
                        loc._ck(lib.lbm2p_comm_unique_id(ident), "lbm2p_comm_unique_id")
                        box[0] = bytes(ident.raw)
                    self.dist.broadcast_object_list(box, src=0)
                    ident.raw = box[0]
            loc._ck(lib.lbm2p_comm_init(ctx, ident, self.part.world, self.part.rank), "lbm2p_comm_init")

    # ---- stepping ------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _stage(self, stage):
        loc = self.local
        return loc._ck(loc._lib.lbm2p_slab_stage(loc._ctx, stage, self._stream()), "lbm2p_slab_stage")

    def _exchange(self, stage):
        self.halo.exchange(stage)

    def step(self):
        self.run(1)

    def run(self, nsteps):
        n = int(nsteps)
        if n <= 0:
            return
        loc = self.local
        if self.transport == "native":
            loc._ck(loc._lib.lbm2p_run_slab(loc._ctx, n, self._stream()), "lbm2p_run_slab")
            self._started = True
            return
        if not self._started:
            self._stage(0)
            self._exchange(0)
            self._started = True
            n -= 1
        for _ in range(n):
            self._stage(1)
            self._exchange(1)
            self._stage(2)
            self._exchange(0)

    # ---- results ---------------------------------------------------------------------------------
    @property
    def launch_count(self):
        return self.local.launch_count

    def local_field(self, name):
        """owned planes of rho / v / F / psi / rho_r / rho_b / solid on this rank"""
        return self.part.owned(getattr(self.local, name).to_numpy())

    def gather_field(self, name, dst=0):
        return _gather_planes(self, self.local_field(name), dst)
