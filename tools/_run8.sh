mkdir -p gpurun_out
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/r2_bench_strong_n$N.json 2> gpurun_out/r2_bench_strong_n$N.err; echo N=$N rc=$?
python tools/bench_brief.py gpurun_out/r2_bench_strong_n$N.json; grep -E "Error" gpurun_out/r2_bench_strong_n$N.err | tail -3
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py > gpurun_out/r2_dist_check_n8.log 2>&1; echo dist_check rc=$?; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2_dist_check_n8.log | tail -6
