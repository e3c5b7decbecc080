mkdir -p gpurun_out
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --no-api-e2e > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo N=$N rc=$?
python tools/bench_brief.py gpurun_out/r2_bench_n$N.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2_bench_n$N.err | tail -3
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py > gpurun_out/r2_dist_check_n8.log 2>&1; echo dist_check rc=$?; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2_dist_check_n8.log | tail -6
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k nccl 2>&1 | tail -3
