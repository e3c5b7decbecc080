mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log | cut -c1-200
(python tools/kbench.py --what fit --genes 5000 --mu-hi 30 --pack product product@g
python tools/kbench.py --what fit --genes 5000 --mu-hi 300 --pack product product@g
python tools/kbench.py --what fit --genes 20000 product product@g
) > gpurun_out/r2h_kbench.log 2>&1
grep -v "^nnz" gpurun_out/r2h_kbench.log | cut -c1-330
timeout 500 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo bench rc=$?; python tools/bench_brief.py gpurun_out/r2h_bench_n1.json 2>&1 | tail -20; tail -3 gpurun_out/r2h_bench_n1.err
