# sparse-kernel experiments: L2 prefetch-size hint x occupancy (rebuilds on the box)
: > gpurun_out/tune.log
for HINT in 0 128 256; do
 for MINB in 6 8; do
  LBM3D_NVCC_FLAGS="-DLBM_SPARSE_L2HINT=$HINT -DLBM_SPARSE_MINB=$MINB" python -m taichi_lbm3d_b200.build --force > /dev/null 2>> gpurun_out/tune.log
  python bench.py --steps 100 --warmup 10 --no-cpu-baseline --sparse --workload porous --size 384 2>> gpurun_out/tune.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HINT=$HINT MINB=$MINB porous384 MLUPS %.0f frac %.4f'%(d['value'],d['roofline']['frac']))"
 done
done
grep -i error gpurun_out/tune.log | head
