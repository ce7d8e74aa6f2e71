"""Prints the few numbers of a bench.py JSON line that matter while tuning."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        o = d["rooflines_other"]
        print(f"{f}: gpus={d['n_gpus']} blocks/s={d['value']:.0f} ms/step={d['ms_per_step']:.2f} e2e_ms={d['e2e']['ms_per_step']:.2f} "
              f"chol_ms={d['roofline']['ms_per_step']:.2f} ({d['roofline']['achieved']:.1f} TF/s, {d['roofline']['frac']:.2f}) "
              f"solve={o['solve_total_ms']:.2f} decode={o['decode']['ms']:.2f} gram={o['gram']['ms']:.2f} classes={[round(x,2) for x in o['chol_class_ms']]}")
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
