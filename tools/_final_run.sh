mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest.log 2>&1; tail -3 gpurun_out/r2_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python tools/bench_brief.py gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
bash tools/_wl_full.sh lrt spline rank 2>&1 | grep -E "rc=|Error"
