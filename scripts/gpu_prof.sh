# ncu captures: launch list + full set for the dense and sparse step kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense -s 5 -c 1 -o gpurun_out/prof_dense python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sparse -s 5 -c 1 -o gpurun_out/prof_sparse_por384 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 384 > gpurun_out/ncu_f2.log 2>&1
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --sparse --workload porous --size 384 > gpurun_out/bench_por384_sparse.json 2> gpurun_out/bench.err
LBM3D_SPARSE_TABLE=full python bench.py --steps 200 --warmup 20 --no-cpu-baseline --sparse --workload porous --size 384 > gpurun_out/bench_por384_sparse_full.json 2>> gpurun_out/bench.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --sparse --workload porous --size 512 > gpurun_out/bench_por512_sparse.json 2>> gpurun_out/bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_por*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s MLUPS %.0f  ms/step %.4f  frac %.4f  nf %d"%(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["fluid_nodes"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench.err
