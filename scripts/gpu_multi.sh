# multi-GPU run: slab tests on one GPU, NCCL parity check, scaling bench N=1,NG
NG=${NG:-2}
(time timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q) > gpurun_out/pytest_slabs.log 2>&1
tail -4 gpurun_out/pytest_slabs.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1
grep multi_gpu_check gpurun_out/multi_check.log || tail -20 gpurun_out/multi_check.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale.err
for n in $(seq 2 $NG); do
 if [ $n = 2 ] || [ $n = 4 ] || [ $n = 8 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/scale_$n.json 2>> gpurun_out/scale.err
 fi
done
python - <<'PY'
import json,glob
base=None
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d["n_gpus"]==1: base=d["value"]
        eff = d["value"]/(d["n_gpus"]*base) if base else 0
        print("%-30s n_gpus %d MLUPS %.0f  ms/step %.4f  per-GPU frac %.4f launches %d  eff %.3f"%(f, d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], eff))
    except Exception as e: print(f, "ERR", e)
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/scale.err | tail -5
