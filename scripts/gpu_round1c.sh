# two-phase parity + timings (colour pass: gather vs tiled), sparse ncu captures, default bench line
(time timeout 1200 python -m pytest tests/test_gpu_two_phase.py -m gpu -x -q) > gpurun_out/pytest_2p.log 2>&1
tail -5 gpurun_out/pytest_2p.log
for T in 0 1; do LBM3D_COLOUR_TILED=$T python scripts/bench_two_phase.py 2>> gpurun_out/bench_2p.err | tee gpurun_out/bench_2p_tiled$T.json | python -c "
import json,sys; [print('tiled $T', d['workload'][:40], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(sys.stdin)]"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sparse_r1b.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 512 > gpurun_out/ncu_sp_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sparse -s 6 -c 1 -o gpurun_out/prof_sparse_r1b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sparse --workload porous --size 512 > gpurun_out/ncu_sp_f.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python bench.py --sparse --workload porous --size 512 --steps 100 > gpurun_out/bench_por512.json 2>> gpurun_out/bench_default.err; cat gpurun_out/bench_por512.json
tail -3 gpurun_out/bench_default.err
