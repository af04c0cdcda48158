#!/usr/bin/env python
"""Benchmark of the fused D3Q19 MRT step (BASELINE.json metric: MLUPS and % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A bench "step" is ONE lattice time step (one fused kernel launch per GPU) over the whole
workload.  N=1 workload: BASELINE config 2, the dense 256^3 lid-driven cavity
(Single_phase/example_cavity.py scaled up; lid vz=0.1 on the x1 face).  N>1: x-slabs of
256 planes per GPU ((256 N) x 256 x 256 cavity, weak scaling), five populations per face
exchanged per step.

value   fluid-node updates per second / 1e6 with the state resident in HBM (CUDA events).
e2e     the same metric for the whole user-level job through the Python class / C ABI with
        HOST buffers: geometry upload (pinned host -> device) + init_simulation + K x step()
        + rho, v and max_v back to the host, all inside the timed region.
roofline  algorithmic bytes (152 B per fluid-node update, BASELINE.json) / kernel time,
        against the measured copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline  the oracle (C/OpenMP restatement of the reference's four-pass step; Taichi
        is not installable here) on the host cores, bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_LUP = 152.0          # 19 fp32 read + 19 fp32 written (BASELINE.json north_star)
SLAB = 256                 # planes per GPU
NY = NZ = 256
LID = [0.0, 0.0, 0.1]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(workload):
    """dram bytes per launch from the committed ncu capture, if one exists for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:  # noqa: BLE001
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(n, steps, warmup, threads=None):
    """The oracle's C/OpenMP four-pass step (stand-in for ti.init(arch=ti.cpu)) on an n^3 cavity."""
    import numpy as np
    from oracle.cref import RefSinglePhaseC
    from taichi_lbm3d_b200.geometry import cavity
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    o = RefSinglePhaseC(n, n, n, kind="fast")
    o.set_solid(cavity(n, n, n))
    o.set_bc_vel(1, LID)
    o.init_simulation()
    nfl = int((o.solid == 0).sum())
    if warmup:
        o.run(warmup)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    return nfl * steps / dt / 1e6, dt, nfl


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # calibrate on a small cube, then pick the largest cube <= 256 that keeps the run bounded
    mlups_cal, _, _ = cpu_reference_run(64, 2, 1)
    budget_s = 150.0
    total_steps = args.steps + args.warmup
    n = 256
    while n > 64 and (n ** 3) * total_steps / (mlups_cal * 1e6) > budget_s:
        n -= 32
    mlups, dt, nfl = cpu_reference_run(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "mlups", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "lid-driven cavity %d^3 dense, D3Q19 MRT single phase (BASELINE config 2 shape)" % n,
                   "note": "Taichi is not installable in this image; this is the oracle's C/OpenMP "
                           "restatement of the reference's four-pass AoS step on the host cores (its strict "
                           "build reproduces the reference source bit for bit, tests/test_reference_pin.py)"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
                         "sample": "%d steps of the %d^3 cavity (%d fluid nodes)" % (args.steps, n, nfl)},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sparse", action="store_true", help="use the compacted fluid-list storage")
    ap.add_argument("--size", type=int, default=SLAB, help="cube edge / planes per GPU (default 256)")
    ap.add_argument("--domain", type=int, default=0,
                    help="fixed GLOBAL cube edge split into x-slabs over the ranks (strong scaling, BASELINE "
                         "config 5: --domain 1024); default 0 = weak scaling with --size planes per GPU")
    ap.add_argument("--workload", default="cavity", choices=["cavity", "porous"],
                    help="cavity: BASELINE config 2 (headline); porous: config 3, periodic sphere pack at "
                         "~20%% porosity, body force fx=1e-6 (use with --sparse)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    from taichi_lbm3d_b200.geometry import cavity, sphere_pack
    n = args.size
    ny = nz = n
    gnx = n * n_gpus
    if args.domain:
        gnx = ny = nz = args.domain
        n = (gnx + n_gpus - 1) // n_gpus
    if os.environ.get("LBM3D_BENCH_SHAPE") and n_gpus == 1:      # tuning experiments: "nx,ny,nz"
        gnx, ny, nz = (int(t) for t in os.environ["LBM3D_BENCH_SHAPE"].split(","))
        n = gnx
    porous = args.workload == "porous"
    if porous:
        r0 = max(3.0, 8.0 * n / 512.0)
        solid = sphere_pack(gnx, ny, nz, 0.80, r0, 2 * r0, seed=n, periodic=True)
    else:
        solid = cavity(gnx, ny, nz)
    nfl_total = int((solid == 0).sum())

    def configure(lb):
        if porous:
            lb.set_force([1e-6, 0.0, 0.0])
        else:
            lb.set_bc_vel_x1(LID)

    if world == 1:
        from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
        lb = LB3D_Solver_Single_Phase(gnx, ny, nz, sparse_storage=args.sparse)
        lb.solid.from_numpy(solid)
        configure(lb)
        lb.init_simulation()
        stepper = lb
    else:
        from taichi_lbm3d_b200.multi_gpu import SlabSolver
        lb = SlabSolver(gnx, ny, nz, sparse_storage=args.sparse)
        lb.set_solid(solid)
        configure(lb)
        lb.init_simulation()
        stepper = lb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    stepper.run(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = stepper.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    stepper.run(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = stepper.launch_count - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    mlups = nfl_total * args.steps / (ms * 1e-3) / 1e6
    max_v = stepper.get_max_v()

    # ---- end to end through the public API with host buffers ---------------------------------
    # the whole user-level job: host geometry in (pinned), init_simulation (table / flag build),
    # K x step(), rho, v and max_v back into pinned host buffers; wall clock between barriers,
    # max over ranks
    del lb, stepper
    torch.cuda.empty_cache()
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, SlabSolver
    own = SlabPartition(gnx, world, rank).own
    pinned = torch.from_numpy(solid).pin_memory()
    rho_pin = torch.empty((own, ny, nz), dtype=torch.float32, pin_memory=True)
    v_pin = torch.empty((own, ny, nz, 3), dtype=torch.float32, pin_memory=True)
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        lb2 = LB3D_Solver_Single_Phase(gnx, ny, nz, sparse_storage=args.sparse)
        lb2.solid.from_numpy(pinned.numpy())            # host geometry in
        configure(lb2)
        lb2.init_simulation()                           # H2D + table build
        t_init = time.perf_counter() - t0
        for _ in range(args.steps):                     # the reference scripts' loop: one call per step
            lb2.step()
        lb2.synchronize()
        t_steps = time.perf_counter() - t0 - t_init
        rho_h = lb2.rho.to_numpy(out=rho_pin.numpy())   # D2H results into pinned host buffers
        v_h = lb2.v.to_numpy(out=v_pin.numpy())
        h2d = solid.nbytes
    else:
        lb2 = SlabSolver(gnx, ny, nz, sparse_storage=args.sparse)
        lb2.set_solid(pinned.numpy())
        configure(lb2)
        lb2.init_simulation()
        t_init = time.perf_counter() - t0
        for _ in range(args.steps):
            lb2.step()
        torch.cuda.synchronize()
        t_steps = time.perf_counter() - t0 - t_init
        rho_pin.numpy()[...] = lb2.local_field("rho")
        v_pin.numpy()[...] = lb2.local_field("v")
        rho_h, v_h = rho_pin.numpy(), v_pin.numpy()
        h2d = (own + 2) * ny * nz
    mv = lb2.get_max_v()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": nfl_total * args.steps / dt / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": h2d * n_gpus / args.steps,
           "d2h_bytes_per_step": (rho_h.nbytes + v_h.nbytes + 4) * n_gpus / args.steps,
           "job": "geometry upload + init_simulation + %d x step() + rho, v, max_v to host%s"
                  % (args.steps, "" if world == 1 else " (every rank its slab; includes creating the NCCL communicator)"),
           "seconds": dt, "seconds_init": t_init, "seconds_steps": t_steps,
           "seconds_readback": dt - t_init - t_steps, "max_v": mv}
    del lb2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    achieved = B_PER_LUP * (nfl_total / n_gpus) * args.steps / (ms * 1e-3) / 1e9    # GB/s per GPU
    workload = "%s%d_%s" % (args.workload, n, "sparse" if args.sparse else "dense")
    wl_text = ("periodic sphere pack %dx%dx%d at porosity %.3f, body force fx=1e-6 (BASELINE config 3 shape)"
               % (gnx, ny, nz, nfl_total / float(gnx * ny * nz))) if porous else \
        ("lid-driven cavity %dx%dx%d, lid vz=0.1 on x1 (BASELINE config 2%s)"
         % (gnx, ny, nz, "" if n_gpus == 1 else ", x-slabs of %d planes per GPU" % n))
    line = {
        "metric": "mlups", "value": mlups, "unit": "MLUPS", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.domain else "weak", "vs_baseline": mlups / 900.0 if n_gpus == 1 else None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s; D3Q19 MRT single phase, %s" % (wl_text, "sparse storage (compact fluid list)"
                                                                    if args.sparse else "dense storage"),
                   "fluid_nodes": nfl_total, "l2_policy": "inputs_exceed_l2 (%.2f GB of populations per GPU vs 126 MB L2)"
                   % (2 * 19 * 4 * float(n) * ny * nz / 1e9),
                   "vs_baseline_note": "900 MLUPS: README.md:5, one A100, grid size unstated",
                   "parallelism": "x-slabs x%d" % n_gpus, "max_v": max_v},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": profiled_traffic(workload), "peak_source": peak_src,
                     "bytes_per_update": B_PER_LUP, "kernel": "k_sparse" if args.sparse else "k_dense",
                     "frac_of_nominal_8TBps": achieved / 8000.0},
        "clocks": clocks, "gpu_launches": launches,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and n_gpus == 1:
        try:
            cores = os.cpu_count() or 1
            # bounded sample: about 12 s of host time on a 128^3 cavity, sized from a 64^3 probe
            cal, _, _ = cpu_reference_run(64, 2, 1)
            nb = 128
            sb = int(max(5, min(400, 12.0 * cal * 1e6 / nb ** 3)))
            v, dtc, nflc = cpu_reference_run(nb, sb, 1)
            line["cpu_baseline"] = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
                                    "sample": "%d steps of a %d^3 cavity (same BCs), C/OpenMP restatement of the "
                                              "reference's 4-pass step (its strict build reproduces the reference "
                                              "source bit for bit, tests/test_reference_pin.py); Taichi unavailable"
                                              % (sb, nb)}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s" % ex}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
