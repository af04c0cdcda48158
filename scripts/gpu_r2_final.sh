# round-2 flow on one GPU, as the driver runs it: the whole GPU suite, smoke, the bench command, the reference arm
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_full.log 2>&1
tail -4 gpurun_out/pytest_full.log
(time timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
cut -c1-400 gpurun_out/bench_r2f.json; tail -2 gpurun_out/bench_r2f.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
cut -c1-300 gpurun_out/bench_ref_n1.json
