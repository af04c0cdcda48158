# what the driver does at round end: gpu tests, smoke, both bench arms
(time timeout 2400 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_full.log 2>&1
tail -6 gpurun_out/pytest_full.log
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1
tail -6 gpurun_out/smoke.log
(time python bench.py --impl reference) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-600; tail -2 gpurun_out/bench_ref.err
(time python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
