for SH in "64,1024,1024" "256,256,256" "128,512,512" "200,131,300"; do
  LBM3D_BENCH_SHAPE=$SH python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>> gpurun_out/shape.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shape $SH  MLUPS %.0f  ms/step %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
done
(time timeout 2400 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_full.log 2>&1
tail -5 gpurun_out/pytest_full.log
tail -3 gpurun_out/shape.err
