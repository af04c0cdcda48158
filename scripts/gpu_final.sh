# round-end flow on one GPU: all GPU tests, smoke, bench line, launch list of the bench command
(time timeout 600 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_full.log 2>&1
tail -5 gpurun_out/pytest_full.log
(time timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-300 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
grep -c k_dense gpurun_out/launches_dense.csv
