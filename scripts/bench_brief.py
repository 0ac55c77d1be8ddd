import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    sh = r.get("kernel_ms_share", {})
    it = d["passes"]["iterations"]
    ms = d["ms_per_step"]
    per = {"dominant": r.get("kernel"), "avg_ms": round(r.get("avg_launch_ms", 0), 3), "launches": r.get("launches"), "frac": round(r.get("frac", 0), 3)}
    print(f"learn {ms:.1f} ms  iters {it}  value {d['value']:.3e}  per-launch ms {per}  recon_err {d['max_abs_coupling_error_vs_truth']:.4f} "
          f"resid {d['max_residual']:.1e} clocks {d.get('clocks')}")
    if d.get("e2e"): print("  e2e", d["e2e"])
    if d.get("cpu_baseline"): print("  cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "seconds")})
