#!/usr/bin/env python
"""Condense gpurun_out/ artefacts of one GPU call into tracked files under profiles/.

    python scripts/summarize_profiles.py <tag> [--launches F.csv] [--rep F.ncu-rep] [--kernel REGEX]
                                               [--copy FILE ...]

  --launches  ncu `--metrics gpu__time_duration.sum` CSV  -> profiles/<tag>_launches.md
              (per-kernel launch count, summed device time, share of the run)
  --rep       ncu `--set full` report -> profiles/<tag>_ncu.md (key raw metrics per launch, stall
              reasons, the most-sampled SASS lines); read here with `ncu -i` (no GPU needed)
  --copy      small text artefacts (bench JSON lines, microbenchmark logs) copied verbatim
"""
from __future__ import annotations

import argparse
import collections
import csv
import hashlib
import io
import json
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches_md(tag: str, path: Path) -> None:
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
        a[2].add(f"{r[gi]}x{r[bi]}")
    total = sum(a[1] for a in agg.values())
    out = [f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
           f"source: `{path.name}`; {len(rows) - 1} launches, {total / 1e6:.3f} ms of device time "
           "(cold-cache, serialised: compare shares, not absolutes)", "",
           "| kernel | launches | total ms | share | grid x block |", "|---|---:|---:|---:|---|"]
    for name, (n, ns, shapes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        sh = ", ".join(sorted(shapes)[:3]) + (" ..." if len(shapes) > 3 else "")
        out.append(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {sh} |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")


def _ncu(rep: Path, page: str) -> list[list[str]]:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", page, "--csv"], check=True, capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def ncu_md(tag: str, rep: Path) -> None:
    raw = _ncu(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = [f"# {tag}: ncu `--set full --clock-control none --import-source on`", "", f"source: `{rep.name}` "
           "(read with `ncu -i ... --page raw|source --csv`)", ""]
    for r in data:
        out += [f"## `{r[ix['Kernel Name']].split('(')[0]}`  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}", "",
                "| metric | value | unit |", "|---|---:|---|"]
        for k in RAW_KEYS:
            if k in ix:
                out.append(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |")
        out += ["", "stall reasons (warps per issue-active cycle):", "", "| reason | ratio |", "|---|---:|"]
        st = []
        for h, i in ix.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, name in sorted(st, reverse=True)[:8]:
            out.append(f"| {name} | {v:.3f} |")
        out.append("")
    src = _ncu(rep, "source")
    # first kernel only: header row is the one starting with "Address"
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    shdr = src[start]
    six = {h: i for i, h in enumerate(shdr)}
    body = []
    for r in src[start + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    tot = sum(int(r[six["# Samples"]]) for r in body)
    out += [f"## most-sampled SASS of the first captured launch ({tot} samples, {len(body)} instructions)", "",
            "| samples | executed | SASS | top stalls |", "|---:|---:|---|---|"]
    stall_cols = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(body, key=lambda r: -int(r[six["# Samples"]]))[:24]:
        stalls = sorted(((int(r[six[c]]), c[6:]) for c in stall_cols if r[six[c]] not in ("0", "")), reverse=True)[:3]
        out.append(f"| {r[six['# Samples']]} | {r[six['Instructions Executed']]} | `{r[six['Source']].strip()}` | "
                   + ", ".join(f"{n}={v}" for v, n in stalls) + " |")
    mix = collections.Counter()
    for r in body:
        op = r[six["Source"]].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        mix[op.split(".")[0]] += int(r[six["Instructions Executed"]])
    tot_i = sum(mix.values())
    out += ["", "executed warp-instruction mix: " + ", ".join(f"{k} {100 * v / tot_i:.1f}%" for k, v in mix.most_common(10))]
    (PROF / f"{tag}_ncu.md").write_text("\n".join(out) + "\n")


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def csrc_digest() -> str:
    """Same digest as bench.py: ties a DRAM-traffic figure to the kernel sources it was captured on."""
    h = hashlib.sha256()
    for p in sorted((ROOT / "epseon_backend_b200" / "csrc").glob("*.cu*")):
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def traffic_json(tag: str, rep: Path, name: str, kernel: str, per_step: int) -> None:
    """profiles/traffic.json[name] = dram__bytes_read + dram__bytes_write per launch of the captured
    launches whose kernel name contains `kernel` (x per_step launches when a bench "launch" is
    several kernel launches, e.g. the 51 chunks of a constant-bank sweep)."""
    raw = _ncu(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, n, names = 0.0, 0, set()
    for r in data:
        if kernel not in r[ix["Kernel Name"]]:
            continue
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[k]].replace(",", "")) * UNIT[units[ix[k]]]
        n += 1
        names.add(r[ix["Kernel Name"]].split("(")[0])
    if n == 0:
        raise SystemExit(f"no captured launch matches {kernel!r}")
    path = PROF / "traffic.json"
    cur = json.loads(path.read_text()) if path.exists() else {}
    cur[name] = {"dram_bytes_per_launch": int(round(tot / n * per_step)), "launches_captured": n, "kernel_launches_per_bench_launch": per_step,
                 "kernel": sorted(names)[0], "source": f"profiles/{tag}_ncu.md ({rep.name}, ncu --set full)", "csrc_digest": csrc_digest()}
    path.write_text(json.dumps(cur, indent=1, sort_keys=True) + "\n")


def traffic_csv(tag: str, path_csv: Path, name: str, per_step: int) -> None:
    """profiles/traffic.json[name] from an APPLICATION-replay ncu CSV (`--replay-mode application
    --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum`): the kernels see the L2
    state the previous launches left, which is what decides whether the constant-bank sweep's carried
    state crosses HBM (kernel replay flushes L2 before every pass and always reads it from DRAM)."""
    rows = list(csv.reader(line for line in open(path_csv) if line.startswith('"')))
    h = rows[0]
    mi, vi, ui, ki = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Kernel Name")
    tot, ids, names = 0.0, set(), set()
    for r in rows[1:]:
        if r[mi] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        tot += float(r[vi].replace(",", "")) * UNIT[r[ui]]
        ids.add(r[h.index("ID")])
        names.add(r[ki].split("(")[0])
    n = len(ids)
    if n == 0:
        raise SystemExit("no launches in " + str(path_csv))
    path = PROF / "traffic.json"
    cur = json.loads(path.read_text()) if path.exists() else {}
    cur[name] = {"dram_bytes_per_launch": int(round(tot / n * per_step)), "launches_captured": n, "kernel_launches_per_bench_launch": per_step,
                 "kernel": sorted(names)[0], "csrc_digest": csrc_digest(),
                 "source": f"profiles/{tag}.csv (ncu --replay-mode application --cache-control none: L2 as the sweep leaves it)"}
    path.write_text(json.dumps(cur, indent=1, sort_keys=True) + "\n")
    shutil.copy(path_csv, PROF / f"{tag}.csv")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", type=Path)
    ap.add_argument("--rep", type=Path)
    ap.add_argument("--copy", type=Path, nargs="*", default=[])
    ap.add_argument("--out", type=Path, help="write into this directory instead of profiles/ (on the GPU box: gpurun_out/profiles)")
    ap.add_argument("--traffic", help="workload key of profiles/traffic.json to (re)write from --rep, e.g. c2 or c5_dform")
    ap.add_argument("--kernel", default="numerov_sweep_kernel", help="kernel-name substring for --traffic")
    ap.add_argument("--per-step", type=int, default=1, help="kernel launches per bench launch (the chunk launches of a c5 sweep)")
    ap.add_argument("--traffic-csv", type=Path, help="application-replay ncu CSV with the dram__bytes metrics -> traffic.json[--traffic]")
    a = ap.parse_args()
    global PROF
    if a.out:
        PROF = a.out
    PROF.mkdir(parents=True, exist_ok=True)
    if a.launches:
        launches_md(a.tag, a.launches)
    if a.rep:
        ncu_md(a.tag, a.rep)
        if a.traffic:
            traffic_json(a.tag, a.rep, a.traffic, a.kernel, a.per_step)
    if a.traffic_csv:
        traffic_csv(a.tag, a.traffic_csv, a.traffic, a.per_step)
    for f in a.copy:
        shutil.copy(f, PROF / f.name)


if __name__ == "__main__":
    main()
