# final captures of round 2 for profiles/: launch list of the driver's bench command, full sets of the two-phase kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --subs none > gpurun_out/ncu_l1.log 2>&1
grep -c "k_dense" gpurun_out/r2_launches_bench.csv
ncu --set full --clock-control none --import-source on -k regex:k2p -s 404 -c 2 -o gpurun_out/r2_prof_two_phase_final python scripts/prof_two_phase.py droplet256 203 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2p -s 404 -c 2 -o gpurun_out/r2_prof_two_phase_sparse_final python scripts/prof_two_phase.py porous384 203 > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out/r2_prof_two_phase*final*.ncu-rep
