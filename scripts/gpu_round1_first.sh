set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
(time python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest.log 2>&1
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
LBM3D_BLOCK=128 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_b128.json 2>> gpurun_out/bench.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --sparse > gpurun_out/bench_sparse.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense -s 5 -c 2 -o gpurun_out/prof_dense_r1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest.log; cat gpurun_out/bench.json gpurun_out/bench_b128.json gpurun_out/bench_sparse.json; tail -3 gpurun_out/bench.err
