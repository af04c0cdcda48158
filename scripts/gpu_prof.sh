# ncu captures for profiles/: launch lists + full sets of the step kernels (1 GPU)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense -s 5 -c 1 -o gpurun_out/prof_dense python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_sparse.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 384 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sparse -s 5 -c 1 -o gpurun_out/prof_sparse python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 384 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2p -s 470 -c 2 -o gpurun_out/prof_2p python scripts/bench_two_phase.py > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep
