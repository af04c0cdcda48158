: > gpurun_out/tune2p.log
for CM in 6 8; do for MM in 4 5 6; do
  LBM3D_NVCC_FLAGS="-DLBM2P_COLOUR_MINB=$CM -DLBM2P_MAIN_MINB=$MM" python -m taichi_lbm3d_b200.build --force > /dev/null 2>> gpurun_out/tune2p.log
  python scripts/bench_two_phase.py 2>> gpurun_out/tune2p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('COLOUR_MINB=$CM MAIN_MINB=$MM', ' | '.join('%s %.0f MLUPS %.3f ms'%(x['workload'][:12],x['mlups'],x['ms_per_step']) for x in d))"
done; done
grep -i "error" gpurun_out/tune2p.log | head -5
