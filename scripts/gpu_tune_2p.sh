for BX in 32 64 128 256; do
  LBM3D_COLOUR_BX=$BX python scripts/bench_two_phase.py 2>> gpurun_out/tune2p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('COLOUR_BX=$BX', ' | '.join('%s %.0f MLUPS %.3f ms'%(x['workload'][:12],x['mlups'],x['ms_per_step']) for x in d))"
done
(time timeout 1200 python -m pytest tests/test_gpu_two_phase.py -x -q) 2>&1 | tail -4
