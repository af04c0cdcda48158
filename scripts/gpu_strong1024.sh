# BASELINE config 5: 1024^3 cavity split into x-slabs (strong scaling points that fit)
for n in 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --domain 1024 --steps 100 --warmup 10 > gpurun_out/strong1024_$n.json 2>> gpurun_out/strong.err
  tail -1 gpurun_out/strong1024_$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d 1024^3: %.0f MLUPS  %.3f ms/step  per-GPU frac %.4f'%(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
grep -v "OMP_NUM\|\*\*\*" gpurun_out/strong.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/weak8.json 2>> gpurun_out/strong.err
tail -1 gpurun_out/weak8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 weak 256-plane slabs: %.0f MLUPS  %.4f ms/step  per-GPU frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['frac']))"
