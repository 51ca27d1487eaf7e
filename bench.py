#!/usr/bin/env python
"""bench.py -- headline benchmark of the Numerov hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|c5] [--impl reference]

Metric: FP64 Numerov grid-steps x trial-energies per second (whole job, all ranks), plus
time-to-all-levels.  One "step" = one complete pass of the hot path over the workload:

  c2 (default; BASELINE.json configs[1]):  Morse (H2-like) potential, 100 000-point grid,
      coarse sweep of 65 536 trial energies, bracketing, k-section refinement of all 17 bound
      levels to 1e-10 relative.  At N GPUs each rank solves its own (slightly perturbed) curve
      -- sharded by potential curve, no data-path collective, weak scaling -- and the located
      levels (17 doubles per rank) are gathered to rank 0 over NCCL.
  c3 (BASELINE.json configs[2]):  tabulated "ab initio" curve (64 knots, natural cubic spline
      resampled to 1 000 000 points), sweep of 4096 trial energies -- the few-energy / long-grid
      regime served by the transfer-matrix scan path; replicas at N > 1 (it does not shard).
  c4 (BASELINE.json configs[3]):  4096 perturbed Morse / LJ curves x 1024 coarse energies each,
      levels 0..7 refined to 1e-10; curves sharded over the ranks (strong scaling).
  c5 (BASELINE.json configs[4]):  dense sweep of 2^24 trial energies on a 200 000-point grid,
      energy-range sharded over the ranks (strong scaling); node-count checksums gathered.

value  = steps executed by all ranks / max-over-ranks CUDA-event time, potentials resident in HBM.
e2e    = same metric through the host-buffer C ABI (eps_set_potentials + eps_solve_levels with
         host pointers: table upload and result download inside the timed region).
roofline.bound = "fp64": the path is FP64-pipe bound; the denominator is
         a DFMA probe measured in this process because MEASURED_PEAKS.json has no FP64 entry.
cpu_baseline / --impl reference: the build's own CPU oracle (the reference repository has no
         implementation of this path -- SURVEY.md section 0), OpenMP over all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from tests import workloads as W  # noqa: E402

METRIC = "numerov_grid_steps_x_trial_energies_per_s"
_REAL_STDOUT = 1  # fd of the run's own stdout (main() moves fd 1 to stderr)
FLOP_PER_STEP = 6  # executed by the 4-instruction X form: 1 DADD + 1 DMUL + 2 DFMA
CPU_SAMPLE_SECONDS = 10.0  # bounded CPU sample of the same workload (cpu_baseline)
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 64 FP64 lanes/SM at the 1965 MHz max clock

# dominant kernel per workload + its DRAM traffic per launch from the committed `ncu --set full`
# captures (dram__bytes_read.sum + dram__bytes_write.sum; profiles/r1g_ncu.md, r1c_scan_ncu.md,
# r1g_c4_ncu.md, r1e_c5_cbank_ncu.md).  Algorithmic bytes = every resident curve's coefficient table
# once per launch (8 B x grid steps x curves).
#   c4: one launch streams the tables of all 4096 curves (measured on the 1-GPU shape).
#   c5: a "launch" of the bench is one sweep = 51 chunk launches of the constant-bank kernel; each
#       carries the per-energy state (28 B read + 28 B written per energy) through HBM by design:
#       881 MB per chunk launch at 2^24 energies = 58 GB/s, under 1 % of the HBM peak.
KERNEL_META = {
    "c2": ("eps::numerov_sweep_kernel<EPT=4,WARPS=4,STRIDE=32> (TMA ring, flat refinement rows)", 839_936, "profiles/r1g_ncu.md"),
    "c3": ("eps::numerov_sweep_kernel<...,SCAN=true> + segment_combine_kernel (transfer-matrix scan)", 8_028_416, "profiles/r1c_scan_ncu.md"),
    "c4": ("eps::numerov_sweep_kernel<EPT=4,WARPS=4,STRIDE=8> (TMA ring, packed refinement rows)", 348_472_832,
           "profiles/r1g_c4_ncu.md (1-GPU shape: 4096 curves per launch)"),
    "c5": ("eps::numerov_cbank_kernel<EPT=4,THREADS=128,STRIDE=32> (constant-bank chunks)", 51 * 881_415_680,
           "profiles/r1e_c5_cbank_ncu.md (1-GPU shape: 51 chunk launches x 881 MB of per-energy state carry)"),
}

C2 = dict(N=100_000, n_coarse=65_536, refine_points=4457, rel_tol=1e-10, max_rounds=8, v_max=16)
C3 = dict(N=1_000_000, nE=4096)
C4 = dict(nC=4096, N=10_000, n_coarse=1024, refine_points=32, rel_tol=1e-10, max_rounds=8, v_max=7)
C5 = dict(N=200_000, nE=1 << 24)


def workload_config(name: str) -> dict:
    """`config` of the JSON line: the same dict for the CUDA arm and for `--impl reference`."""
    l2 = "flushed between timed steps (256 MiB memset)"
    if name == "c2":
        return {"workload": "c2: Morse (H2-like) 100k-point grid, 65536 trial energies coarse sweep + k-section "
                            "refinement of all 17 bound levels to 1e-10 rel.; one curve per GPU (curve-sharded)",
                "grid_points": C2["N"], "trial_energies_coarse": C2["n_coarse"],
                "refine_points_per_level": C2["refine_points"], "levels": C2["v_max"] + 1, "l2": l2}
    if name == "c3":
        return {"workload": "c3: tabulated curve (64 knots, natural cubic spline) resampled to a 1M-point grid, sweep of "
                            "4096 trial energies through the transfer-matrix scan path; replicas at N > 1",
                "grid_points": C3["N"], "trial_energies": C3["nE"], "l2": l2}
    if name == "c4":
        return {"workload": "c4: 4096 perturbed Morse/LJ curves x 1024 coarse energies, levels 0..7 refined to 1e-10 "
                            "rel.; curves sharded over the ranks",
                "curves": C4["nC"], "grid_points": C4["N"], "trial_energies_coarse": C4["n_coarse"],
                "refine_points_per_level": C4["refine_points"], "levels": C4["v_max"] + 1, "l2": l2}
    return {"workload": "c5: dense sweep of 2^24 trial energies on a 200k-point grid, energy-range sharded",
            "grid_points": C5["N"], "trial_energies": C5["nE"], "l2": l2}


def rank_curve(rank: int):
    """C2 curve of a rank: the H2-like Morse curve, parameters perturbed +-2 % for rank > 0."""
    rng = np.random.default_rng(7000 + rank)
    j = 1.0 + (0.02 * (2.0 * rng.random(3) - 1.0) if rank else np.zeros(3))
    De, a, re = W.H2["De"] * j[0], W.H2["a"] * j[1], W.H2["re"] * j[2]
    rmin, rmax, N = 0.2, 12.0, C2["N"]
    V = W.morse(De, re, a, rmin, rmax, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(rmin, rmax, N))
    return V, s, 0.0, De - 1.0, (De, a)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 20 ms from a
    thread (nvidia_ml_py), falling back to `nvidia-smi -lms` when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc, self.nvml, self._stop = device, [], None, None, False

    def _nvml_loop(self):
        n, h = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        while not self._stop:
            try:
                clk = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((time.time(), float(clk), pw, [k for k, b in bits.items() if mask & b]))
            except Exception:  # noqa: BLE001 - sampling must never break the benchmark
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as n

            n.nvmlInit()
            idx = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[self.device])
            h = n.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            threading.Thread(target=self._nvml_loop, daemon=True).start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.nvml is not None:
            self._stop = True
            time.sleep(0.02)
            sel = [r for r in self.rows if t0 <= r[0] <= t1]
            reasons = sorted({x for r in sel for x in r[3]})
            return {"sm_mhz": float(np.median([r[1] for r in sel])) if sel else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(sel), "power_w_max": max((r[2] for r in sel), default=None), "reasons": reasons,
                    "source": "nvml, 20 ms period, inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                mhz.append(clk)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "samples": len(mhz),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def host_threads() -> int:
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1; the CPU legs run on
    rank 0 only and are meant to use the whole box)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, torch, world, rank, local
    return None, None, 1, 0, 0


def cpu_solve_c2(orc, V, s, E_lo, E_hi):
    F, *_ = orc.prep(V, s)
    t = time.perf_counter()
    lev, wid, nb, rounds, steps = orc.solve_levels(F, s, E_lo, E_hi, C2["n_coarse"], 0, C2["v_max"],
                                                   C2["refine_points"], C2["rel_tol"], C2["max_rounds"])
    return time.perf_counter() - t, steps, lev


def run_reference(args) -> None:
    """--impl reference: the CPU implementation of the path on the host cores.  The reference
    repository has none, so this times the build's oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import Oracle

    orc = Oracle(omp=True, threads=host_threads())
    if args.workload == "c5":
        w = W.c5(C5["N"], C5["nE"])
        F, *_ = orc.prep(w["V"], w["s"])
        n_steps = F.size
        nE = 1 << 16  # bounded sample of the 2^24-energy sweep; work is exactly linear in nE
        dE = (w["E_hi"] - w["E_lo"]) / (C5["nE"] - 1)

        def one():
            t = time.perf_counter()
            orc.sweep_uniform(F, w["s"], w["E_lo"], dE * (C5["nE"] // nE), 0, nE, tails=False)
            return time.perf_counter() - t, n_steps * nE

        sample = f"2^16 of the 2^24 energies (every 256th) on the 200k grid, {orc.threads} threads"
        cfg = workload_config("c5")
    elif args.workload == "c3":
        w = W.c3(C3["N"], C3["nE"])
        F, *_ = orc.prep(w["V"], w["s"])
        dE = (w["E_hi"] - w["E_lo"]) / (C3["nE"] - 1)

        def one():
            t = time.perf_counter()
            orc.sweep_uniform(F, w["s"], w["E_lo"], dE, 0, C3["nE"], tails=False)
            return time.perf_counter() - t, F.size * C3["nE"]

        sample = f"the full C3 sweep (4096 energies x 1M-point grid), {orc.threads} threads"
        cfg = workload_config("c3")
    elif args.workload == "c4":
        w = W.c4(64, C4["N"], C4["n_coarse"])

        def one():
            t, steps = time.perf_counter(), 0
            for c in range(64):
                F, *_ = orc.prep(w["V"][c], w["s"])
                steps += orc.solve_levels(F, w["s"], w["E_lo"][c], w["E_hi"][c], C4["n_coarse"], 0, C4["v_max"],
                                          C4["refine_points"], C4["rel_tol"], C4["max_rounds"])[4]
            return time.perf_counter() - t, steps

        sample = f"64 of the 4096 curves (full level solve each), {orc.threads} threads"
        cfg = workload_config("c4")
    else:
        V, s, E_lo, E_hi, _ = rank_curve(0)

        def one():
            dt, steps, _ = cpu_solve_c2(orc, V, s, E_lo, E_hi)
            return dt, steps

        sample = f"full C2 solve (coarse 65536 + refinement of 17 levels), {orc.threads} threads"
        cfg = workload_config("c2")
    for _ in range(args.warmup):
        one()
    tot_t = tot_s = 0.0
    for _ in range(args.steps):
        dt, st = one()
        tot_t += dt
        tot_s += st
    val = tot_s / tot_t
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.workload in ("c4", "c5") else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": orc.threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference repo has no implementation of this path; this is the build's CPU oracle (port)",
    })


def emit(line: dict) -> None:
    """The ONE JSON line of the run, written to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main() -> None:
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner,
    # compiler chatter of build()) goes to stderr instead.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["c2", "c3", "c4", "c5"], default="c2")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import __graft_entry__ as ge

    dist, torch, world, rank, local = dist_setup(args.gpus)
    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    from epseon_backend_b200 import cabi

    ctx = cabi.Context(local)
    sampler = ClockSampler(local)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def pinned(a: np.ndarray) -> np.ndarray:
        """The step's input table in page-locked host memory (the e2e leg copies it host -> device
        inside the timed region, every step)."""
        out = ctx.pinned_empty(np.atleast_2d(a).shape, np.float64)
        out[...] = a
        return out

    if args.workload == "c2":
        V, s, E_lo, E_hi, (De, a) = rank_curve(rank)
        V = pinned(V)
        ctx.set_potentials(V, s)
        n_steps = ctx.curve_info(0).n_steps

        def step_resident():
            return ctx.solve_levels(E_lo, E_hi, C2["n_coarse"], 0, C2["v_max"], C2["refine_points"],
                                    C2["rel_tol"], C2["max_rounds"])

        def step_e2e():
            ctx.set_potentials(V, s)  # host table -> prep -> H2D
            return step_resident()  # ... -> levels, widths, counts D2H

        cfg = workload_config("c2")
        scaling = "weak"
    elif args.workload == "c3":
        w = W.c3(C3["N"], C3["nE"])
        V, s = pinned(w["V"]), w["s"]
        ctx.set_potentials(V, s)
        n_steps = ctx.curve_info(0).n_steps
        dE = (w["E_hi"] - w["E_lo"]) / (C3["nE"] - 1)

        def step_resident():
            return ctx.sweep_uniform(w["E_lo"], w["E_hi"], C3["nE"], nodes=False, tails=False)

        def step_e2e():
            ctx.set_potentials(V, s)
            n, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], C3["nE"], nodes=True, tails=False)
            return n

        cfg = workload_config("c3")
        scaling = "weak"
    elif args.workload == "c4":
        w = W.c4(C4["nC"], C4["N"], C4["n_coarse"])
        per = C4["nC"] // world
        sl = slice(rank * per, (rank + 1) * per)
        V, s = pinned(w["V"][sl]), w["s"]
        E_lo4, E_hi4 = np.ascontiguousarray(w["E_lo"][sl]), np.ascontiguousarray(w["E_hi"][sl])
        ctx.set_potentials(V, s)
        n_steps = ctx.curve_info(0).n_steps

        def step_resident():
            return ctx.solve_levels(E_lo4, E_hi4, C4["n_coarse"], 0, C4["v_max"], C4["refine_points"], C4["rel_tol"],
                                    C4["max_rounds"])

        def step_e2e():
            ctx.set_potentials(V, s)
            return step_resident()

        cfg = workload_config("c4")
        scaling = "strong"
    else:
        w = W.c5(C5["N"], C5["nE"])
        ctx.set_potentials(w["V"], w["s"])
        n_steps = ctx.curve_info(0).n_steps
        from epseon_backend_b200 import multi

        sl = multi.curve_shard(C5["nE"], world, rank)  # contiguous slice of the global energy grid
        j0, per = sl.start, sl.stop - sl.start
        dE = float(multi.global_step(w["E_lo"], w["E_hi"], C5["nE"]))
        V, s = pinned(w["V"]), w["s"]

        def step_resident():
            return ctx.sweep_grid(w["E_lo"], dE, j0, per, nodes=False, tails=False)

        def step_e2e():
            ctx.set_potentials(V, s)
            n, _, _ = ctx.sweep_grid(w["E_lo"], dE, j0, per, nodes=True, tails=False)  # 4 B/energy D2H
            return n

        cfg = workload_config("c5")
        scaling = "strong"

    gather_bufs = {}

    def gather_small(arr: np.ndarray):
        """Gather a small per-rank result (the only inter-GPU traffic of the path): pinned staging,
        one all-gather over NCCL, one copy back; every rank ends up holding all ranks' rows."""
        if dist is None:
            return [arr]
        arr = np.ascontiguousarray(arr)
        key = (arr.shape, arr.dtype.str)
        if key not in gather_bufs:
            h_in = torch.from_numpy(np.empty_like(arr)).pin_memory()
            d_in = torch.empty_like(h_in, device="cuda")
            d_out = torch.empty((world,) + tuple(arr.shape), dtype=h_in.dtype, device="cuda")
            h_out = torch.empty(d_out.shape, dtype=h_in.dtype).pin_memory()
            gather_bufs[key] = (h_in, d_in, d_out, h_out)
        h_in, d_in, d_out, h_out = gather_bufs[key]
        h_in.copy_(torch.from_numpy(arr))
        d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [h_out[r].numpy().copy() for r in range(world)] if rank == 0 else None

    def result_digest(res) -> np.ndarray:
        if args.workload == "c2":
            return res[0][0]  # 17 level energies
        if args.workload == "c4":
            return res[0]  # [curves of this rank][8] level energies
        return np.zeros(1)

    # ---- warm-up (incl. the NCCL communicator behind the result gather), FP64 probe ----
    for _ in range(args.warmup):
        res = step_resident()
        gather_small(result_digest(res))
    ctx.sync()
    fp64_peak, _ = ctx.fp64_probe()

    def timed_run(step_fn, k: int):
        """K steps; CUDA events on the ctx stream around each step; L2 flushed between steps."""
        ms_total, results = 0.0, None
        for _ in range(k):
            ctx.l2_flush()
            barrier()
            ctx.timer_start()
            results = step_fn()
            digest = gather_small(result_digest(results))
            ms = ctx.timer_stop()
            if dist is not None:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            ms_total += ms
        return ms_total, results, digest

    # ---- timed: resident ----
    barrier()
    ctx.stats_reset()
    sampler.start()
    t_wall0 = time.time()
    ms_res, res, digest = timed_run(step_resident, args.steps)
    t_wall1 = time.time()
    st = ctx.stats()
    clocks = sampler.stop(t_wall0, t_wall1)
    steps_rank = float(st.grid_steps)
    if dist is not None:
        t = torch.tensor([steps_rank], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        steps_all = float(t.item())
    else:
        steps_all = steps_rank
    value = steps_all / (ms_res * 1e-3)
    sweep_rate = steps_rank / (st.sweep_ms * 1e-3)  # this rank's dominant kernel, averaged over its launches
    launches = int(st.sweep_launches + st.other_launches)

    # ---- timed: end-to-end through the host-buffer C ABI ----
    for _ in range(2):
        step_e2e()
    barrier()
    ctx.stats_reset()
    ms_e2e, _, _ = timed_run(step_e2e, args.steps)
    st2 = ctx.stats()
    steps2 = float(st2.grid_steps)
    if dist is not None:
        t = torch.tensor([steps2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        steps2 = float(t.item())
    e2e_value = steps2 / (ms_e2e * 1e-3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "time_to_all_levels_ms": ms_res / args.steps if args.workload in ("c2", "c4") else None,
            "e2e": {"value": e2e_value, "unit": "steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(st2.h2d_bytes // args.steps),
                    "d2h_bytes_per_step": int(st2.d2h_bytes // args.steps)},
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64", "kernel": KERNEL_META[args.workload][0],
                "achieved": FLOP_PER_STEP * sweep_rate / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": FLOP_PER_STEP * sweep_rate / 1e12 / fp64_peak,
                "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": NOMINAL_FP64_TFLOPS,
                "frac_of_nominal": FLOP_PER_STEP * sweep_rate / 1e12 / NOMINAL_FP64_TFLOPS,
                "flop_per_step": FLOP_PER_STEP, "steps_per_s_kernel": sweep_rate,
                "fp64_instr_per_step": 4, "sweep_launches": int(st.sweep_launches),
                "avg_launch_ms": st.sweep_ms / max(1, st.sweep_launches),
                "traffic": KERNEL_META[args.workload][1], "traffic_source": KERNEL_META[args.workload][2],
                "algorithmic_bytes_per_launch": 8 * int(n_steps) * int(ctx.n_curves),
            },
            "clocks": clocks,
        }
        if args.workload == "c2":
            exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
            lev0 = digest[0]
            line["levels_found"] = int(np.sum(np.isfinite(lev0)))
            line["max_rel_err_vs_analytic_rank0"] = float(np.max(np.abs(lev0 - exact) / exact))
        if not args.no_cpu_baseline:
            from oracle import Oracle

            orc = Oracle(omp=True, threads=host_threads())

            def cpu_timed(fn, min_seconds=CPU_SAMPLE_SECONDS):
                """Repeat fn() -> (steps, result) until min_seconds of CPU work; -> (steps/s, seconds, reps, first result)."""
                tot_t, tot_s, reps, first = 0.0, 0.0, 0, None
                while tot_t < min_seconds:
                    t0 = time.perf_counter()
                    st_c, r = fn()
                    tot_t += time.perf_counter() - t0
                    tot_s += st_c
                    reps += 1
                    first = r if first is None else first
                return tot_s / tot_t, tot_t, reps, first

            if args.workload == "c2":
                Vc, sc, El, Eh, _ = rank_curve(0)

                def one():
                    _, csteps, clev = cpu_solve_c2(orc, Vc, sc, El, Eh)
                    return csteps, clev

                rate, secs, reps, clev = cpu_timed(one)
                line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": orc.threads, "kind": "port",
                                        "sample": f"the full C2 solve (coarse 65536 + refinement), repeated {reps}x, OpenMP oracle",
                                        "seconds": secs,
                                        "levels_bit_identical_to_gpu": bool(np.array_equal(
                                            clev.view(np.uint64), digest[0].view(np.uint64)))}
            elif args.workload == "c3":
                F, *_ = orc.prep(V[0], s)

                def one():
                    n_c, _, _ = orc.sweep_uniform(F, s, w["E_lo"], dE, 0, C3["nE"], tails=False)
                    return F.size * C3["nE"], n_c

                rate, secs, reps, n_cpu = cpu_timed(one)
                n_gpu, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], C3["nE"], nodes=True, tails=False)
                line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": orc.threads,
                                        "kind": "port", "sample": f"the full C3 sweep, repeated {reps}x, OpenMP oracle",
                                        "seconds": secs,
                                        "nodes_bit_identical_to_gpu": bool(np.array_equal(n_cpu, n_gpu[0]))}
                line["scan"] = {"launches": ctx.counter(ctx.CNT_SCAN_LAUNCHES), "flagged": ctx.counter(ctx.CNT_SCAN_FLAGGED)}
            elif args.workload == "c4":
                n_sample = 64

                def one():
                    csteps, same = 0, True
                    for c in range(n_sample):
                        F, *_ = orc.prep(V[c], s)
                        lv, _, _, _, st_c = orc.solve_levels(F, s, E_lo4[c], E_hi4[c], C4["n_coarse"], 0, C4["v_max"],
                                                             C4["refine_points"], C4["rel_tol"], C4["max_rounds"])
                        csteps += st_c
                        same &= bool(np.array_equal(lv.view(np.uint64), digest[0][c].view(np.uint64)))
                    return csteps, same

                rate, secs, reps, same = cpu_timed(one)
                line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": orc.threads, "kind": "port",
                                        "sample": f"{n_sample} of the 4096 curves (full level solve each), repeated {reps}x, OpenMP oracle",
                                        "seconds": secs, "levels_bit_identical_to_gpu": same}
            else:
                F, *_ = orc.prep(V[0], s)
                nE = 1 << 18

                def one():
                    n_c, _, _ = orc.sweep_uniform(F, s, w["E_lo"], dE * (C5["nE"] // nE), 0, nE, tails=False)
                    return F.size * nE, n_c

                rate, secs, reps, _ = cpu_timed(one)
                line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": orc.threads,
                                        "kind": "port", "sample": f"2^18 of the 2^24 energies (every 64th), repeated {reps}x",
                                        "seconds": secs}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
