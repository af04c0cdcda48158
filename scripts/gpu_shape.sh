for SH in "200,131,300" "131,131,131" "100,300,700"; do
  LBM3D_BENCH_SHAPE=$SH python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>> gpurun_out/shape.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shape $SH  MLUPS %.0f  ms/step %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
done
tail -3 gpurun_out/shape.err
bash scripts/gpu_prof.sh
