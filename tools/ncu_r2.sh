# round-2 profile of the headline step (one GPU): launch list + one --set full capture of the step's kernels
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-api-e2e --no-cpu-baseline"
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dx_ -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:dx_ingest_tma|dx_fit_lane|dx_bh_fused|dx_pack_overflow" --launch-skip 8 -c 4 -o gpurun_out/r2_step -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out
