"""One line per bench JSON file: python tools/bench_summary.py files..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    r = d.get("roofline") or {}
    fit = d.get("gp_fit_ms") or {}
    ex = (d.get("stages") or {}).get("fit_exchange") or {}
    print(f"{f}: n_gpus {d.get('n_gpus')} {d.get('scaling')} value {d['value']:.4g} ms/step {d['ms_per_step']:.2f} wall {d.get('wall_ms_per_step', 0):.2f} "
          f"e2e {d['e2e']['value']:.4g} | roof frac {r.get('frac', 0):.3f} launch {r.get('avg_launch_ms', 0):.3f} ms | fit K+chol {fit.get('k_build_plus_cholesky', 0):.2f} "
          f"inv {fit.get('inversion_for_predict', 0):.2f} exch {ex.get('ms', 0) or 0:.2f} ms ({(ex.get('achieved_gbs') or 0):.0f} GB/s) wall {fit.get('fit_wall_ms', 0):.1f} | "
          f"clk {d.get('clocks', {}).get('sm_mhz')} | idx {d.get('result', {}).get('global_index')} best {d.get('result', {}).get('best')}")
    if "cpu_baseline" in d:
        print("    cpu", round(d["cpu_baseline"]["value"], 1), "cores", d["cpu_baseline"]["cores"])
