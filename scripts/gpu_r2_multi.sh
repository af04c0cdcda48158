# round-2 multi-GPU flow (NG GPUs): parity checks of both solvers, the driver's bench command
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_n${NG}.log 2>&1
grep multi_gpu_check gpurun_out/multi_check_n${NG}.log || tail -20 gpurun_out/multi_check_n${NG}.log
timeout 600 $TR --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_n${NG}_final.json 2> gpurun_out/bench_n${NG}_final.err
cut -c1-300 gpurun_out/bench_n${NG}_final.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_n${NG}_final.err | tail -3
timeout 300 $TR --master-port 29512 scripts/multi_gpu_check_2p.py > gpurun_out/multi_check_2p_n${NG}.log 2>&1
grep -i "multi_gpu_check" gpurun_out/multi_check_2p_n${NG}.log | tail -8 || tail -20 gpurun_out/multi_check_2p_n${NG}.log
