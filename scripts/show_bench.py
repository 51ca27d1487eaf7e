#!/usr/bin/env python
"""Print the headline and the sub-records of a bench.py JSON line in one table."""
import json
import sys

d = json.load(open(sys.argv[1]))


def brief(name, r):
    par = {k: v for k, v in r.get("cpu_baseline", {}).items() if "identical" in k}
    par.update({k: v for k, v in r.items() if "identical" in k or k == "max_rel_diff_vs_ksection"})
    rf = r.get("roofline", {})
    print(f"{name:14s} value {r['value']:.3e}  ms/step {r['ms_per_step']:9.3f}  e2e {r['e2e']['ms_per_step']:9.3f} ms  "
          f"frac {rf.get('frac', float('nan')):.3f}  launches {r.get('gpu_launches')}  {par}")


print("n_gpus", d["n_gpus"], "steps", d["steps"], "clocks", d.get("clocks"))
brief("c5", d)
if "accurate_mode" in d:
    brief("c5 accurate", d["accurate_mode"])
for k, r in d.get("sub_records", {}).items():
    brief(k, r)
    for m in ("accurate_mode", "cooley_mode"):
        if m in r:
            brief(f"{k} {m[:-5]}", r[m])
