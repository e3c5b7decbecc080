import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    k = d.get("kernels", {})
    print("value %.3g genes/s  ms/step %.3f | ingest %.3f ms (%.1f%% of HBM peak)  fit %.3f ms  wald %.4f ms | e2e %.3g genes/s | niter %.2f conv %.3f | clocks %s" % (
        d["value"], d["ms_per_step"], k["dx_ingest_tma_kernel"]["ms"], 100 * d["roofline"]["frac"],
        k["dx_fit_grouped_kernel"]["ms"], k["dx_wald_from_fit_kernel"]["ms"], d["e2e"]["value"],
        d["config"]["mean_niter"], d["config"]["frac_converged"], d["clocks"]))
    if d.get("cpu_baseline"):
        print("cpu_baseline", d["cpu_baseline"])
