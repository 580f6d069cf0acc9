#!/usr/bin/env python
"""Benchmark of the fp64 cubed-sphere PPM advection step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config 4 of BASELINE.json): N=1536 cells per panel edge, Nair-Lauritzen
divergent flow (vf=3), Gaussian hill, par-default scheme PPM-PL07 / RK1 / SP-AVLT /
ET-DG / MT-0 / MF-PR, dt = 0.00625*16/N.  A step = adv_time_step + update_adv.
metric = cell-updates/s = 6 N^2 K / time.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TUPLE = (3, 1, 1, 3, 1, 3)      # recon, dp, opsplit, et, mt, mf  (par/advection.par defaults)
VF, IC = 3, 2
DT16 = 0.00625                  # canonical dt(N) = DT16 * 16 / N (src/advection_error.py:43-58)
GOLDEN, GOLDEN_KS = "config_N1536_vf3.npz", (5, 20)
NAME = "config4: N=%d vf=3 divergent flow, Gaussian hill, PPM-PL07/RK1/SP-AVLT/ET-DG/MT-0/MF-PR"
BYTES_PER_CELL = 40.0           # Q r/w, two winds, sqrt(g)


def select_config(c):
    """--config 4 (default, the headline): BASELINE.json config 4.  --config 3: config 3 -- N=768, Nair-Lauritzen
    non-divergent flow (time-dependent wind), RK2 departure points, AVLT-RK2-DG-PR; recorded, not the headline."""
    global TUPLE, VF, DT16, GOLDEN, GOLDEN_KS, NAME
    if c == 3:
        TUPLE, VF, DT16 = (3, 2, 1, 3, 1, 3), 2, 0.0125
        GOLDEN, GOLDEN_KS = "config_N768_vf2.npz", (10, 50)
        NAME = "config3: N=%d vf=2 Nair-Lauritzen non-divergent flow, Gaussian hill, PPM-PL07/RK2/SP-AVLT/ET-DG/MT-0/MF-PR"


def workload(N):
    return {"workload": NAME % N, "N": N, "cells": 6 * N * N, "dt": DT16 * 16 / N,
            "l2": "inputs (%d MB/step) exceed the 126 MB L2" % round(BYTES_PER_CELL * 6 * N * N / 1e6)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a moment to start: do not open the timed region before it samples."""
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:   # region shorter than the sampling period: take the nearest samples
            for ts, line in sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:3]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- CPU arm
def oracle_state(g, N):
    from oracle import step as ost
    sim = ost.Simulation(g, DT16 * 16 / N, 5, IC, VF, 1, *TUPLE)
    ost.init_vars_adv(g, sim)
    return sim, ost


def cpu_steps(g, sim, ost, k0, n):
    t = time.perf_counter()
    ost.run(g, sim, n, k0)
    return time.perf_counter() - t


def run_reference(args):
    """--impl reference: the reference's own algorithm (numpy oracle port, bit-identical to
    the reference under the numpy shim) timed on the host cores.  /root/reference is pure
    Python and does not exist on the GPU box, so there is no oracle/_ref build.
    Every one of the K + W steps is a full advection step of the same scheme and wind; the
    sample is bounded through the grid size: N = 1536 (the workload itself, ~20 s per step on
    one core) when K + W <= 8, else N = 768 (a quarter of the cells per step) so that the run
    ends within a few minutes.  --cpu-n overrides.  cell-updates/s is size-normalised."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.cpu_n if args.cpu_n else (1536 if args.steps + args.warmup <= 8 else 768)
    from oracle.grid import LeanGrid
    g = LeanGrid(N)
    sim, ost = oracle_state(g, N)
    k = 0
    for _ in range(args.warmup):
        cpu_steps(g, sim, ost, k, 1)
        k += 1
    t = 0.0
    for _ in range(args.steps):
        t += cpu_steps(g, sim, ost, k, 1)
        k += 1
    value = 6.0 * N * N * args.steps / t
    cfg = workload(args.n)            # the arm's config is the GPU arm's, key for key; what ran is in cpu_sample
    line = {"impl": "reference", "cpu_sample": "ran at N=%d (%d cells per step), same scheme, wind and dt rule" % (N, 6 * N * N), "metric": "cell-updates/s (fp64 advection step)", "value": value,
            "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": "port",
                             "sample": "%d full steps (after %d warm-up) of the numpy oracle at N=%d, same scheme and "
                                       "wind as the workload (whole-array numpy is single threaded; host has %d cores)"
                                       % (args.steps, args.warmup, N, os.cpu_count())},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import ctypes as C
    import pycs_b200  # noqa: F401
    from pycs_b200 import cs_datastruct, advection_ic, advection_vars, advection_timestep, parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = args.n
    cfg = workload(N)
    dt = cfg["dt"]
    t_setup = time.time()
    # lean grid: geometry, wind, initial condition and diagnostics are generated on the device (csrc/grid.cu);
    # --host-grid builds the reference's numpy arrays instead (10-15 s at N=1536)
    g = cs_datastruct.cubed_sphere(N, lean=not args.host_grid)
    sim = advection_ic.adv_simulation_par(g, dt, 5, IC, VF, 1, *TUPLE, device=local)
    advection_vars.init_vars_adv(g, sim)
    dev = sim.dev
    if not dev.fused_supported():
        raise RuntimeError("fused kernel unavailable")
    slab = (4, N + 4)
    if world > 1:
        # strong scaling: the rows of every panel are split over the ranks (csrc/mgpu.cu)
        slab = parallel.shard(sim)
    setup_s = time.time() - t_setup
    sm, name = dev.sm_count()
    Q0 = np.asarray(sim.Q).copy()

    def barrier():
        dev.call("pycs_synchronize")
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of exactly this path (sharded or not) against the reference's own numbers: the fixture
    # tests/golden/config_N1536_vf3.npz holds Q of the unmodified reference on a 51 x 51 x 6 sample after
    # 5 and 20 steps of this workload; every rank checks the sample rows of its slab.
    parity = None
    gp = os.path.join(ROOT, "tests", "golden", GOLDEN)
    if os.path.exists(gp) and not args.quick and int(np.load(gp)["N"]) == N:
        ref = np.load(gp)
        idx = ref["sample_index"]
        rows = [n for n, i in enumerate(idx) if slab[0] <= i < slab[1]]
        worst, kdone = 0.0, 0
        for kk in GOLDEN_KS:
            dev.call("pycs_run", kdone, kk - kdone, 1)
            kdone = kk
            Qs = np.asarray(sim.Q)[np.ix_(idx[rows], idx, np.arange(6))]
            want = ref["Q_k%d" % kk]
            worst = max(worst, float(np.max(np.abs(Qs - want[rows])) / np.max(np.abs(want))) if rows else 0.0)
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([worst], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            worst = float(tt.item())
        parity = {"relerr": worst, "against": "tests/golden/%s (unmodified reference, Q on a 51x51x6 sample at "
                  "k = %d and %d)" % ((GOLDEN,) + GOLDEN_KS), "tolerance": 1e-12, "ok": bool(worst <= 1e-12)}
        sim.Q[...] = Q0                   # back to the initial state for the timed run
        barrier()

    # ---- device-resident run: W warm-up steps, then K timed steps (CUDA events on the handle's stream).
    # The clock sampler starts before the warm-up and a barrier sits immediately before the timed
    # region, so that the ranks enter it together (a rank that is late would be waited for inside the
    # others' timed kernels).
    sampler = ClockSampler(local)
    sampler.start()
    dev.call("pycs_run", 0, args.warmup, 1)
    sampler.wait_first()
    l0 = dev.launches()
    ms = C.c_float()
    barrier()
    t0 = time.time()
    dev.call("pycs_run_timed", args.warmup, args.steps, 1, C.byref(ms))
    t1 = time.time()
    barrier()
    clocks = sampler.stop(t0, t1)
    launches = dev.launches() - l0
    ms_total = float(ms.value)
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    cells = 6.0 * N * N                      # whole sphere: the ranks share one problem
    own_cells = 6.0 * N * (slab[1] - slab[0])
    value = cells * args.steps / (ms_total * 1e-3)
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "steps": args.steps, "ms_per_step": ms_total / args.steps,
                              "value": value, "launches": int(launches), "clocks": clocks,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("PYCS_")}}))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    # ---- roofline: the step kernel alone, CUDA events around back-to-back launches.
    # kernel_ms: best of 3 bursts of 50 launches (the HBM peak it is compared with is a burst copy,
    # best of 10); kernel_ms_sustained: average over a long run (the 1 kW power cap shows there).
    kms = C.c_float()
    sep = 1 if (VF == 3 and TUPLE[1] == 1) else 0      # the flavour the timed steps launch
    dev.call("pycs_time_step_kernel", 5, sep, C.byref(kms))
    bursts = []
    for _ in range(3):
        dev.call("pycs_time_step_kernel", 50, sep, C.byref(kms))
        bursts.append(float(kms.value) / 50)
    k_ms = min(bursts)
    reps = max(20, args.steps)
    dev.call("pycs_time_step_kernel", reps, sep, C.byref(kms))
    k_ms_sustained = float(kms.value) / reps
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([k_ms, k_ms_sustained], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        k_ms, k_ms_sustained = float(tt[0].item()), float(tt[1].item())
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_CELL * own_cells / (k_ms * 1e-3) / 1e9
    tb, rows, nblk = C.c_int32(), C.c_int32(), C.c_int32()
    dev.call("pycs_step_kernel_info", C.byref(tb), C.byref(rows), C.byref(nblk))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # one ncu --set full capture of the same kernel at the single-GPU size (profiles/)
            traffic = tj.get("dram_bytes_per_launch") if world == 1 else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dev.step_kernel_name(),
                "kernel_ms": k_ms, "kernel_ms_sustained": k_ms_sustained,
                "frac_sustained": BYTES_PER_CELL * own_cells / (k_ms_sustained * 1e-3) / 1e9 / peak,
                "timing": "CUDA events on the launch stream; kernel_ms = best of 3 bursts of 50 back-to-back launches, "
                          "kernel_ms_sustained = average of %d" % reps,
                "algorithmic_bytes_per_launch": BYTES_PER_CELL * own_cells,
                "per": "GPU (each rank updates %d of %d rows of every panel)" % (slab[1] - slab[0], N),
                "peak_source": peak_src, "grid": {"ctas": nblk.value, "threads": tb.value, "rows_per_chunk": rows.value},
                "share_of_step": k_ms_sustained / (ms_total / args.steps)}

    # ---- end to end: adv_time_step through host buffers (pinned), H2D + step + D2H every step
    sim.Q[...] = Q0
    hostQ = torch.empty(Q0.shape, dtype=torch.float64, pin_memory=True)
    hq = hostQ.numpy()
    hq[...] = Q0
    ptr = hq.ctypes.data_as(C.POINTER(C.c_double))
    e2e_steps = max(3, min(args.steps, 10))
    for k in range(1, 3):
        dev.call("pycs_adv_time_step_host", ptr, k, k * dt, 1)
    barrier()
    te = time.perf_counter()
    for k in range(3, 3 + e2e_steps):
        dev.call("pycs_adv_time_step_host", ptr, k, k * dt, 1)
    barrier()
    te = time.perf_counter() - te
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    # a sharded handle moves only its slab: rows [row_lo, row_hi) of the (P, P, 6) array, both ways
    slab_bytes = int(Q0.nbytes) if world == 1 else (slab[1] - slab[0]) * Q0.shape[1] * Q0.shape[2] * 8
    moved = slab_bytes
    if world > 1:
        tt = torch.tensor([float(slab_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        moved = int(tt.item())
    e2e = {"value": cells * e2e_steps / te, "unit": "cell-updates/s", "h2d_bytes_per_step": moved,
           "d2h_bytes_per_step": moved, "ms_per_step": 1e3 * te / e2e_steps, "steps": e2e_steps,
           "api": "pycs_adv_time_step_host (adv_time_step + update_adv on a host numpy Q; every rank moves its own "
                  "rows of Q over its own PCIe link)"}

    # ---- CPU baseline (rank 0, N=1 only): the numpy oracle on the same grid
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        Nc = args.cpu_n if args.cpu_n else 768
        from oracle.grid import LeanGrid
        og = LeanGrid(Nc)
        osim, ost = oracle_state(og, Nc)
        cpu_steps(og, osim, ost, 0, 1)
        nst = args.cpu_steps
        tc = cpu_steps(og, osim, ost, 1, nst)
        cpu = {"value": 6.0 * Nc * Nc * nst / tc, "unit": "cell-updates/s", "cores": 1, "kind": "port",
               "sample": "%d steps of the numpy oracle at N=%d, same scheme and wind (%.1f s; numpy whole-array "
                         "ops are single threaded, host has %d cores)" % (nst, Nc, tc, os.cpu_count())}

    if rank == 0:
        line = {"metric": "cell-updates/s (fp64 advection step)", "value": value, "unit": "cell-updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "parity": parity, "parity_relerr": parity["relerr"] if parity else None,
                "gpu_launches": int(launches), "clocks": clocks, "device": name, "sm_count": sm,
                "setup_s": setup_s, "grid": "host numpy" if args.host_grid else "device-generated (lean)",
                "parallelism": "single GPU" if world == 1 else
                "%d row slabs per panel, peer-mapped halo stores over NVLink (csrc/mgpu.cu)" % world,
                "wind_path": ("separable (U(t)=U(0)cos(pi t/T) scaled in-kernel" if sep else
                              "basis winds (the wind of a step combined from static basis fields by one kernel per direction") +
                             "; the exposed wind arrays are caught up lazily when something reads them)"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", type=int, default=0, help="cells per panel edge (0: the configuration's own, 1536 / 768)")
    ap.add_argument("--config", type=int, default=4, choices=(3, 4), help="BASELINE.json configuration (4 = headline)")
    ap.add_argument("--cpu-n", type=int, default=0,
                    help="N of the CPU sample (0: reference arm 1536 if steps + warmup <= 8 else 768; cpu_baseline leg 768)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="steps of the cpu_baseline sample (N=768: ~1.7 s each on one core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--host-grid", action="store_true", help="build the grid with host numpy like the reference")
    ap.add_argument("--quick", action="store_true", help="tuning runs: only the device-resident timed steps (no roofline, "
                    "e2e, CPU or parity legs); the line is not a bench record")
    args = ap.parse_args()
    select_config(args.config)
    if not args.n:
        args.n = 1536 if args.config == 4 else 768
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()
