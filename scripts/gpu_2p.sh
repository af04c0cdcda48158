(time timeout 1200 python -m pytest tests/test_gpu_two_phase.py -x -q) > gpurun_out/pytest_2p.log 2>&1
tail -12 gpurun_out/pytest_2p.log
python scripts/bench_two_phase.py > gpurun_out/bench_2p.json 2> gpurun_out/bench_2p.err; cat gpurun_out/bench_2p.json; tail -3 gpurun_out/bench_2p.err
