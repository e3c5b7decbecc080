mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
python tools/bh_time.py 2>&1 | tail -5
DXB200_BH_ONECTA=1 python tools/bh_time.py 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo bench rc=$?; python tools/bench_brief.py gpurun_out/r2d_bench_n1.json 2>&1 | tail -20; tail -3 gpurun_out/r2d_bench_n1.err
