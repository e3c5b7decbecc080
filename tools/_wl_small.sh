for w in rank lrt spline; do
  case $w in rank) a="--cells 200000 --genes 2000";; lrt) a="--cells 20000 --genes 2000";; spline) a="--cells 20000 --genes 1000";; esac
  timeout 600 python bench.py --workload $w $a --steps 2 --warmup 1 > gpurun_out/wl_small_$w.json 2> gpurun_out/wl_small_$w.err; echo "$w rc=$?"
  tail -3 gpurun_out/wl_small_$w.err; cut -c1-1500 gpurun_out/wl_small_$w.json
done
