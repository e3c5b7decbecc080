"""One-paragraph summary of a bench.py JSON line (stdin or a file argument)."""
import sys, json
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for l in src:
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    k = d.get("kernels", {}) or {}
    r = d.get("roofline", {}) or {}
    print("N=%s value %.4g %s  ms/step %.4f  scaling %s" % (d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], d.get("scaling")))
    print("  kernels: " + "  ".join("%s %.4f ms" % (kk.split("(")[0], vv["ms"]) for kk, vv in k.items()))
    if r:
        print("  roofline: %s frac %.3f (%.0f GB/s), whole step frac %.3f" % (r.get("kernel"), r.get("frac", 0), r.get("achieved", 0),
                                                                             (r.get("whole_step") or {}).get("frac", 0)))
    for key in ("e2e", "e2e_cabi"):
        e = d.get(key)
        if e:
            print("  %s: %.4g genes/s, %.1f ms per call, matches %s" % (key, e["value"], e.get("ms_per_call", 0), e.get("matches_device_path")))
    c = d.get("config", {})
    print("  niter %.2f conv %.3f exchange_ok %s launch %s | clocks %s" % (c.get("mean_niter", 0), c.get("frac_converged", 0),
                                                                        c.get("exchange_matches_nccl"), c.get("launch"), d.get("clocks")))
    if d.get("per_rank_ms"):
        for i, pr in enumerate(d["per_rank_ms"]):
            print("  rank %d: %s" % (i, pr))
    if d.get("cpu_baseline"):
        print("  cpu_baseline", d["cpu_baseline"])
