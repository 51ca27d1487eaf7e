#!/bin/bash
# HBM traffic of the constant-bank sweep's chunk launches with the L2 state a real sweep leaves behind:
# application replay, no cache flush between kernels (ncu's default kernel replay flushes L2 before every
# pass, so it always sees the carried state come from DRAM).  60 mid-sweep launches per group size.
for g in 2 3 4; do
  EPS_CB_GROUP=$g ncu --replay-mode application --cache-control none --clock-control none \
     --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:numerov_cbank \
     --launch-skip 120 -c 60 --csv --log-file /tmp/cbt_$g.csv python scripts/ncu_target.py c5 0 1 > /tmp/cbt_$g.log 2>&1
  python - $g <<'PY'
import csv, sys
g = sys.argv[1]
rows = list(csv.reader(l for l in open(f"/tmp/cbt_{g}.csv") if l.startswith('"')))
h = rows[0]; mi, vi, ui = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
    tot.setdefault(r[mi], []).append(v)
n = len(tot["dram__bytes_read.sum"])
rd, wr, t = (sum(tot[k]) / n for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
print(f"group {g}: {n} launches, per launch: DRAM read {rd/1e6:.2f} MB, write {wr/1e6:.2f} MB, duration {t/1e6:.3f} ms")
PY
done
