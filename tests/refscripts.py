"""Running the reference's own CASE SCRIPTS (Single_phase/example_*.py) against a module of the
name they import.

The scripts are the drop-in contract of SURVEY 8(b): they `import taichi as ti`, call `ti.init`,
`import LBM_3D_SinglePhase_Solver as lb3dsp`, build the solver, set it up and loop over `step()`.
`run_script` executes a script's text VERBATIM with
  * a no-op `taichi` module (the two Taichi lines then do nothing),
  * `LBM_3D_SinglePhase_Solver` bound to the module given by the caller -- the product
    (examples/LBM_3D_SinglePhase_Solver.py) or the recorder below,
  * `range` capped so that the script's 2000 .. 150000-step loop ends after `max_iter` iterations.

/root/reference does not exist on the GPU box, and reference sources are not copied into this
repository.  What IS committed (tests/golden/ref_example_traces.json, written by
tests/golden/make_example_traces.py) is the sequence of API calls each script makes, recorded here
by running the script verbatim against `Recorder`: the GPU test replays that sequence on the product
class, and a CPU test re-records it from /root/reference and checks that nothing changed.
"""
import builtins
import os
import sys
import types

import numpy as np


def _taichi_stub():
    ti = types.ModuleType("taichi")
    ti.cpu, ti.gpu, ti.cuda = "cpu", "gpu", "cuda"
    ti.f32, ti.i32 = np.float32, np.int32
    ti.init = lambda *a, **k: None
    return ti


def run_script(text, solver_module, max_iter, filename="<reference example>"):
    """exec the script text verbatim; returns its globals"""
    real_range = builtins.range

    def capped_range(*a):
        r = real_range(*a)
        return r[:max_iter] if len(r) > max_iter else r

    saved = {k: sys.modules.get(k) for k in ("taichi", "LBM_3D_SinglePhase_Solver")}
    sys.modules["taichi"] = _taichi_stub()
    sys.modules["LBM_3D_SinglePhase_Solver"] = solver_module
    ns = {"__name__": "__main__", "range": capped_range, "print": lambda *a, **k: None}
    try:
        exec(compile(text, filename, "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns


class _RecField:
    def __init__(self, rec, name):
        self._rec, self._name = rec, name

    def from_numpy(self, arr):
        a = np.asarray(arr)
        self._rec.calls.append([self._name + ".from_numpy", {"shape": list(a.shape), "solid_index":
                                np.flatnonzero(a.reshape(-1) > 0).tolist()}])


class Recorder:
    """stands in for LB3D_Solver_Single_Phase and writes down what a script does with it"""
    last = None

    def __init__(self, *args, **kwargs):
        self.calls = [["__init__", {"args": list(args), "kwargs": kwargs}]]
        self.steps = 0
        self.solid = _RecField(self, "solid")
        self.fx = self.fy = self.fz = 0.0
        Recorder.last = self

    def step(self):
        self.steps += 1
        if not self.calls or self.calls[-1][0] != "step":
            self.calls.append(["step", 0])
        self.calls[-1][1] += 1

    def get_max_v(self):
        self.calls.append(["get_max_v", None])
        return 0.0

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def call(*args):
            self.calls.append([name, [a if not isinstance(a, np.ndarray) else a.tolist() for a in args]])
        return call


def recorder_module():
    m = types.ModuleType("LBM_3D_SinglePhase_Solver")
    m.LB3D_Solver_Single_Phase = Recorder
    return m


def record(path, max_iter):
    """the API-call sequence of the reference script at `path` (run verbatim, loop capped)"""
    with open(path) as fh:
        text = fh.read()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))
    try:
        run_script(text, recorder_module(), max_iter, path)
    finally:
        os.chdir(cwd)
    return Recorder.last.calls


def replay(calls, cls):
    """the recorded sequence on a real solver class; returns the solver"""
    lb = None
    for name, arg in calls:
        if name == "__init__":
            lb = cls(*arg["args"], **arg["kwargs"])
        elif name == "solid.from_numpy":
            g = np.zeros(int(np.prod(arg["shape"])))
            g[arg["solid_index"]] = 1
            lb.solid.from_numpy(g.reshape(arg["shape"]))
        elif name == "step":
            for _ in range(arg):
                lb.step()
        elif name == "get_max_v":
            lb.get_max_v()
        else:
            getattr(lb, name)(*arg)
    return lb
