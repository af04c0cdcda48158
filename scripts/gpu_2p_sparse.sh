(time timeout 1200 python -m pytest tests/test_gpu_two_phase.py tests/test_gpu_two_phase_slabs.py -m gpu -x -q) > gpurun_out/pytest_2p.log 2>&1
tail -12 gpurun_out/pytest_2p.log
python scripts/bench_two_phase.py > gpurun_out/bench_2p.json 2> gpurun_out/bench_2p.err; python -c "
import json; [print(d['workload'][:70], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(open('gpurun_out/bench_2p.json'))]"; tail -3 gpurun_out/bench_2p.err
