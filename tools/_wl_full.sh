mkdir -p gpurun_out
for w in "$@"; do
  timeout 1200 python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "$w rc=$?"
  grep -v "Warning\|warn\|design_scale=" gpurun_out/r2_bench_$w.err | tail -3; cut -c1-200 gpurun_out/r2_bench_$w.json
done
