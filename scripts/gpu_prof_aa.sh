# ncu launch list + full capture of the in-place (AA) sparse step pair, porous 512^3 (1 GPU)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sparse_aa.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 512 > gpurun_out/ncu_aa_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sparse -s 6 -c 2 -o gpurun_out/prof_sparse_aa python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 512 > gpurun_out/ncu_aa_f.log 2>&1
ls -la gpurun_out/*.ncu-rep
