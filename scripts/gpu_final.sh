# round-end flow on one GPU: all GPU tests, smoke, bench line, two-phase bench, two-phase sparse ncu
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_full.log 2>&1
tail -5 gpurun_out/pytest_full.log
(time timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1
tail -5 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-400 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 200 python scripts/bench_two_phase.py > gpurun_out/bench_2p.json 2> gpurun_out/bench_2p.err; python -c "
import json; [print(d['workload'][:74], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(open('gpurun_out/bench_2p.json'))]"; tail -3 gpurun_out/bench_2p.err
cat > /tmp/p2.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase
from taichi_lbm3d_b200.geometry import sphere_pack
n = 384
solid = sphere_pack(n, n, n, 0.80, 6.0, 12.0, seed=n, periodic=True)
psi = np.ones(solid.shape, np.float32); psi[:n // 4] = -1.0
lb = LB3D_Solver_Two_Phase(n, n, n, sparse_storage=True)
lb.solid.from_numpy(solid); lb.psi.from_numpy(psi)
lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7
lb.init_simulation(); lb.run(14); lb.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2p -s 16 -c 2 -o gpurun_out/prof_2p_sparse python /tmp/p2.py > gpurun_out/ncu_2ps.log 2>&1
ls -la gpurun_out/prof_2p_sparse.ncu-rep
