# new sparse tests + two-phase parity with the tiled colour pass + two-phase timings (tiled on / off)
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or sparse_tables or strict_bit_identical or degenerate or full_size") > gpurun_out/pytest_sparse_b.log 2>&1
tail -3 gpurun_out/pytest_sparse_b.log
(time timeout 1200 python -m pytest tests/test_gpu_two_phase.py -m gpu -x -q) > gpurun_out/pytest_2p.log 2>&1
tail -5 gpurun_out/pytest_2p.log
python scripts/bench_two_phase.py > gpurun_out/bench_2p_tiled.json 2> gpurun_out/bench_2p.err; python -c "
import json; [print(d['workload'][:40], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(open('gpurun_out/bench_2p_tiled.json'))]"
LBM3D_COLOUR_TILED=0 python scripts/bench_two_phase.py > gpurun_out/bench_2p_gather.json 2>> gpurun_out/bench_2p.err; python -c "
import json; [print(d['workload'][:40], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(open('gpurun_out/bench_2p_gather.json'))]"
for XS in 8 16 64; do LBM3D_COLOUR_XSEG=$XS python scripts/bench_two_phase.py 2>> gpurun_out/bench_2p.err | python -c "
import json,sys; [print('xseg $XS', d['workload'][:40], '%.0f MLUPS %.4f ms frac %.3f'%(d['mlups'], d['ms_per_step'], d['roofline']['frac'])) for d in json.load(sys.stdin)]"; done
tail -3 gpurun_out/bench_2p.err
python scripts/sparse_ab.py 512 ab::LBM3D_AA=0 aa::LBM3D_AA=1 ab_h256:h256:LBM3D_AA=0 2>&1 | tee gpurun_out/sparse_ab2.log
