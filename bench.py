#!/usr/bin/env python
"""bench.py -- GR2M cell-months/s (+ routed cell-months/s) on N B200s, beside the host-CPU oracle.

One "step" = one pass of the calibration hot path over the whole workload: every candidate
parameter set simulated on every subbasin for every month, outlet series and objective value per set
(Optim_GR2MSemiDistr's OFUN, R/Optim_GR2MSemiDistr.R:125-217, batched).  Workload = BASELINE.json
configs[3] "C4": 1e5 subbasins x 1e4 parameter sets x 480 months, synthetic forcing.
With N GPUs the parameter sets are sharded across ranks (forcing replicated, no data-path
collective); objective values are all-gathered over NCCL at the end of each step.  Total work is
fixed, so scaling is "strong".

  value     cell-months/s with forcing and parameters resident in HBM, objective left on the device
  e2e       same metric through the host-buffer C ABI (gr2m_basin_create + gr2m_objective_batch +
            destroy): forcing H2D from pinned memory and objective D2H inside the timed region
  roofline  recursion kernel vs the FP64 pipe (measured DFMA peak)
  routing   secondary metric: routed cell-months/s of the batched WFA sweep on a synthetic D8 grid
  cpu_baseline  the CPU oracle (a port: the reference's R/airGR/TauDEM path cannot run here) timed
            on this box's host cores on a bounded sample of the same workload

`--impl reference` times that CPU path alone (all host threads), same metric/config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gr2m_cell_months_per_sec"
UNIT = "cell-months/s"
FLOP_PER_UNIT_ALGO = 67          # nominal FP64 operations per (cell, set, month)
# ncu counters of the shipped recursion kernel (profiles/r2_calib_counters.json, written by scripts/ncu_summary.py from
# the capture named inside): instructions per unit, FP64-pipe utilisation, DRAM bytes per unit
try:
    NCU = json.load(open(os.path.join(ROOT, "profiles", "r2_calib_counters.json")))
except Exception:
    NCU = None


def pipe_model(units, kernel_ms, clocks, sms=148):
    """Cycles per warp and unit on one SM sub-partition against what the FP64 pipe alone needs (a warp-wide FP64
    instruction occupies the 16-lane pipe for 2 cycles)."""
    try:
        mhz = float(clocks["sm_mhz"]) if clocks and clocks.get("sm_mhz") else 1965.0
        took = kernel_ms * 1e-3 * mhz * 1e6 * sms * 4 / (units / 32.0)
        need = 2.0 * NCU["fp64_pipe_inst_per_unit"]
        return {"cycles_taken_per_warp_unit": took, "fp64_pipe_cycles_needed": need, "fp64_pipe_busy_live": need / took}
    except Exception as ex:
        return {"error": repr(ex)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=100_000)
    ap.add_argument("--sets", type=int, default=10_000)
    ap.add_argument("--months", type=int, default=480)
    ap.add_argument("--route-n", type=int, default=20_000, help="side of the synthetic D8 grid (0 = skip routing)")
    ap.add_argument("--route-batch", type=int, default=32, help="months per routing batch")
    ap.add_argument("--route-flags", type=int, default=0, help="WFA_* flags for the routing sweep")
    ap.add_argument("--route-terrain", default="both", choices=["tilted", "filled", "both"],
                    help="tilted: plane + roughness below the tilt (no depressions, round-1 grid); filled: SURVEY 8(d) terrain, "
                         "plane + 5 octaves of value noise with slopes comparable to the tilt, depressions filled and D8 derived "
                         "by wfa_flowdir_from_dem on the device")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-routed", action="store_true", help="skip the routed-calibration figure")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--literal", action="store_true", help="use the literal (IEEE division) kernel variant")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def workload_name(a):
    return f"C4 calibration sweep: {a.cells} subbasins x {a.sets} parameter sets x {a.months} months"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower() == "active"})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def cpu_port(a, nthreads, seconds):
    """Time the CPU oracle's OFUN batch on a bounded sample (same forcing generator, same months)."""
    import oracle
    from gr2msemidistr_b200 import synth
    oracle.build()
    cells = min(a.cells, 2000)
    P, E, area, nd = synth.forcing(cells, a.months)
    region = np.zeros(cells, np.int32)
    qobs = synth.qobs_from(np.full(a.months, 100.0))
    # calibrate the sample size on a tiny run, then size it for ~`seconds`
    par = synth.param_sets(4 * max(nthreads, 8))
    t0 = time.perf_counter()
    oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, warmup=36, nthreads=nthreads, want_qt=False)
    dt = time.perf_counter() - t0
    rate = cells * a.months * len(par) / dt
    nsets = int(max(len(par), min(a.sets, rate * seconds / (cells * a.months))))
    nsets = max(nthreads, (nsets // nthreads) * nthreads)
    par = synth.param_sets(nsets)
    t0 = time.perf_counter()
    oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, warmup=36, nthreads=nthreads, want_qt=False)
    dt = time.perf_counter() - t0
    units = cells * a.months * nsets
    return {"value": units / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{cells} subbasins x {nsets} sets x {a.months} months of the C4 generator, {dt:.1f} s, "
                      f"oracle/gr2m_oracle.c (gcc -O2) with {nthreads} pthreads over parameter sets"}, units, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    vals, last = [], None
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    for i in range(a.warmup + a.steps):
        cb, units, dt = cpu_port(a, nthreads, per_step)
        if i >= a.warmup:
            vals.append((units, dt))
        last = cb
    units = sum(u for u, _ in vals)
    dt = sum(d for _, d in vals)
    v = units / dt
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / max(1, a.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "inputs larger than L2"},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's own R/airGR/TauDEM path cannot run in this image (no R, gfortran, MPI, TauDEM); "
                    "this is the CPU oracle port, a lower bound on the reference's time"}
    print(json.dumps(line), flush=True)


def cpu_routing(a):
    """CPU oracle AreaD8 on a bounded dense sample, one month per host thread (TauDEM's own
    parallelism is 8 MPI ranks per month, R/Routing_GR2MSemiDistr.R:117)."""
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from gr2msemidistr_b200 import synth
    oracle.build()
    n = min(a.route_n, 3000)
    nthreads = os.cpu_count() or 1
    d8 = synth.d8_grid(n, n).numpy()
    rng = np.random.default_rng(7)
    W = [rng.uniform(0, 100, (n, n)) for _ in range(nthreads)]
    oracle.aread8(d8, W[0], contcheck=False)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(nthreads) as ex:
        list(ex.map(lambda w: oracle.aread8(d8, w, contcheck=False), W))
    dt = time.perf_counter() - t0
    return {"value": n * n * nthreads / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{n} x {n} grid of the same generator x {nthreads} months, {dt:.1f} s, orc_aread8 (Kahn queue, FP64), "
                      f"one month per thread"}


def bench_routed_calibration(a, capi, torch, n=6000, nb=100, nsets=2048):
    """north_star's whole loop at a gauge: simulate every candidate set, route, score the routed discharge of one polygon
    (Run -> Routing -> hydroGOF), through the host-buffer call a calibration driver makes
    (gr2m_objective_routed_batch: parameters up, objectives down).  Geometry: one convergent drainage tree
    (synth.d8_convergent) cut into nb x nb subbasin blocks; the gauge is the outlet block, which every subbasin drains
    through, as at a gauged watershed.  `value` counts the subbasins upstream of the gauge (the others would not enter
    its objective and are not simulated)."""
    from gr2msemidistr_b200 import synth
    months = a.months
    dev = torch.device("cuda", torch.cuda.current_device())
    d8 = synth.d8_convergent(n, device=dev)
    t0 = time.perf_counter()
    plan = capi.Plan(dev_ptr=d8.data_ptr(), shape=(n, n))
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    src, off, cells = synth.block_polygons(n, n, nb, nb, margin=max(1, n // nb // 4))
    router = plan.router(src, off, cells)
    t_geo = time.perf_counter() - t0
    nsub = src.shape[0]
    P, E, area, nd = synth.forcing(nsub, months, seed=4321)
    par = synth.param_sets(nsets, seed=4322)
    qobs = synth.qobs_from(np.full(months, 5.0e4))
    gauge = nsub - 1
    nk, owner = router.gauge_upstream(gauge)
    nup = int((owner >= 0).sum())
    out = {}
    with capi.Basin(P, E, area, nd, np.zeros(nsub, np.int32), 1) as b:
        for label, env in (("fused", None), ("materialised", "1")):
            if env:
                os.environ["GR2M_ROUTED_MATERIALISE"] = env
            ns = nsets if not env else min(nsets, 256)
            try:
                b.objective_routed_batch(router, par[:32], qobs, gauge, warmup=36, which=capi.OF_KGE)
                dt, kern_ms, all_s = 1e30, 0.0, []
                for _ in range(3):
                    t0 = time.perf_counter()
                    got = b.objective_routed_batch(router, par[:ns], qobs, gauge, warmup=36, which=capi.OF_KGE)
                    all_s.append(time.perf_counter() - t0)
                    if all_s[-1] < dt:
                        dt, kern_ms = all_s[-1], b.last_kernel_ms()
            finally:
                os.environ.pop("GR2M_ROUTED_MATERIALISE", None)
            out[label] = {"sets": ns, "seconds": dt, "seconds_each_call": all_s, "device_ms": kern_ms, "sets_per_s": ns / dt,
                          "value": float(nup) * months * ns / dt, "finite_objectives": int(np.isfinite(got["of"]).sum())}
    info = router.info()
    router.close()
    plan.close()
    f = out["fused"]
    return {"metric": "routed_calibration_cell_months_per_sec",
            "workload": f"{nsub} subbasins ({nb} x {nb} blocks of a convergent {n} x {n} D8 tree) x {nsets} parameter sets x "
                        f"{months} months, KGE of the routed discharge at the outlet polygon: {nup} subbasins drain through "
                        f"it ({nk} maximal candidate node(s))",
            "value": f["value"], "unit": UNIT,
            "value_def": "upstream subbasins x months x sets / wall time of gr2m_objective_routed_batch (host parameters in, "
                         "objectives out); subbasins that do not reach the gauge are not simulated and not counted",
            "sets_per_s": f["sets_per_s"], "upstream_subbasins": nup, "maximal_candidates": nk,
            "plan_and_operator_build_s": t_geo, "operator": info, **out,
            "materialised_note": "same call with GR2M_ROUTED_MATERIALISE=1 on 256 sets: QS of all subbasins written, operator "
                                 "applied to sets x months columns, gauge row read (the path used for several regions or more "
                                 "than four maximal candidates)",
            "api": "gr2m_objective_routed_batch"}


def bench_router(a, capi, torch, plan, n, months, dense_ms_one_gpu):
    """Routing as Routing_GR2MSemiDistr uses it (one value per subbasin, per-polygon maximum): the
    month-invariant operator built once for the geometry, then applied to all months."""
    from gr2msemidistr_b200 import synth
    nb = max(1, min(100, n // 40))
    src, off, cells = synth.block_polygons(n, n, nb, nb, margin=max(1, n // nb // 4))
    nsub = src.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()
    router = plan.router(src, off, cells)
    t_build = time.perf_counter() - t0
    QS = torch.rand((nsub, months), dtype=torch.float64, device=dev) * 300.0
    QR = torch.empty_like(QS)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        router.apply_dev(QS.data_ptr(), QR.data_ptr(), months)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    info = router.info()
    del QS, QR
    router.close()
    return {"workload": f"{nsub} subbasins ({nb} x {nb} blocks) on the same {n} x {n} grid x {months} months, QS resident",
            "build_s": t_build, "apply_ms": best, "dense_sweep_ms_same_months": dense_ms_one_gpu,
            "interior_cells": int(off[-1]), **info,
            "note": "same QR as the dense sweep, bit for bit (tests/test_gpu_router.py); build includes the H2D of the "
                    "polygon cell lists and is paid once per geometry"}


def bench_routing(a, capi, torch, rank, world, dist, terrain="tilted"):
    """Secondary metric: batched WFA sweep, months sharded across ranks (plan replicated)."""
    from gr2msemidistr_b200 import synth
    n = a.route_n
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()
    lake_share = None
    if terrain == "filled":
        z, valid = synth.fractal_dem(n, n, device=dev)
        v8 = valid.to(torch.uint8).contiguous()
        del valid
        d8 = torch.empty((n, n), dtype=torch.int16, device=dev)
        filled = torch.empty((n, n), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        capi.flowdir_from_dem_dev(z.data_ptr(), v8.data_ptr(), n, n, d8.data_ptr(), filled.data_ptr())
        lake_share = float(((filled > z + 1e-9) & (v8 != 0)).double().mean().item())
        del z, v8, filled
    else:
        d8 = synth.d8_grid(n, n, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = capi.Plan(dev_ptr=d8.data_ptr(), shape=(n, n))
    t_plan = time.perf_counter() - t0
    del d8
    torch.cuda.empty_cache()
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    ncell, ld = n * n, a.route_batch
    months_total = a.months
    my_months = months_total // world + (1 if rank < months_total % world else 0)
    nb = (my_months + ld - 1) // ld
    A = torch.empty((ncell, ld), dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def one_pass(timed):
        tot = 0.0
        for b in range(nb):
            A.uniform_(0.0, 100.0, generator=gen)       # synthetic dense weights, made on the device (untimed)
            m = min(ld, my_months - b * ld)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.accumulate_dev(A.data_ptr(), m, ld, a.route_flags)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot

    one_pass(False)
    ms = one_pass(True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    units = float(ncell) * months_total
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    if os.path.exists(peaks_file):
        hbm, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    gbs = units / world * 16.0 / (ms * 1e-3) / 1e9
    # DRAM bytes of one 32-month sweep of the 20000^2 grid, summed over its launches (ncu, profiles/r2_route_traffic.txt):
    # 202.7 GB on the tilted grid, 239.2 GB on the filled one, against 204.8 GB algorithmic; scaled to this rank's months
    per_sweep = {"tilted": 202.7e9, "filled": 239.2e9}[terrain]
    traffic = per_sweep * my_months / 32.0 if n == 20000 else None
    real_gbs = traffic / (ms * 1e-3) / 1e9 if traffic else None
    out = {"metric": "routed_cell_months_per_sec", "value": units / (ms * 1e-3), "unit": UNIT,
           "workload": f"C5 WFA: {n} x {n} synthetic D8 grid x {months_total} months, dense FP64 weights, "
                       f"{ld}-month batches, months sharded over {world} GPU(s)",
           "terrain": ("plane tilted to the south-east + 5 octaves of value noise with slopes comparable to the tilt, closed "
                       f"depressions filled ({100 * lake_share:.1f} % of the cells raised) and D8 derived on the device "
                       "(wfa_flowdir_from_dem_dev; SURVEY 8(d))") if terrain == "filled" else
                      "plane tilted to the south-east + roughness below the tilt (no depressions to fill; round-1 grid)",
           "ms_per_pass": ms, "levels": plan.nlevels, "cells_in_order": plan.norder,
           "plan_build_s": t_plan, "grid_gen_s": t_gen,
           "value_incl_plan_build": units / (ms * 1e-3 + t_plan),
           "value_incl_plan_build_note": "the D8 plan (levels, reaches, stream offsets) is built once per geometry and reused "
                                         "by every sweep; this figure charges it to this one pass over the months",
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                        "traffic": traffic, "algorithmic_bytes_per_unit": 16, "peak_source": hbm_src,
                        "traffic_def": "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of this rank's "
                                       "sweeps (a 32-month sweep is 250-600 launches: chain-segment levels + reach gather + "
                                       "reach sweep), ncu on one sweep of the same grid scaled by months "
                                       "(profiles/r2_route_traffic.txt)",
                        "achieved_real_bytes": real_gbs, "frac_real_bytes": real_gbs / hbm if real_gbs else None,
                        "traffic_note": "tilted grid: 202.7 GB moved per sweep = the algorithmic 204.8 GB (source cells cost "
                                        "8 B, the chain segments keep the re-read of a predecessor row out of DRAM); filled grid "
                                        "(few sources, sheet flow): 239.2 GB = 1.17 x algorithmic, i.e. the sweep runs at "
                                        "frac_real_bytes of the copy bandwidth in bytes actually moved"}}
    out["basins"] = plan.basin_info()
    out["reaches"] = plan.reach_info()
    del A
    torch.cuda.empty_cache()
    if rank == 0 and terrain == "tilted":
        try:
            out["operator"] = bench_router(a, capi, torch, plan, n, months_total, ms * world)
        except Exception as ex:       # secondary figure: report, do not lose the sweep number
            out["operator"] = {"error": repr(ex)}
    plan.close()
    if rank == 0 and not a.no_cpu and terrain == "tilted":
        out["cpu_baseline"] = cpu_routing(a)
    return out


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from gr2msemidistr_b200 import capi, synth
    from gr2msemidistr_b200 import dist as gdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        g.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    capi.lib()
    capi.set_device(local)
    torch.cuda.set_device(local)

    # ---- inputs (every rank generates the same seeded arrays; forcing is replicated) ----------
    P, E, area, nd = synth.forcing(a.cells, a.months)
    region = np.zeros(a.cells, np.int32)
    par_all = synth.param_sets(a.sets)
    qobs = synth.qobs_from(np.full(a.months, 100.0) * (1 + 0.5 * np.sin(np.arange(a.months) / 6.0)))
    s0 = a.sets * rank // world
    s1 = a.sets * (rank + 1) // world
    par = np.ascontiguousarray(par_all[s0:s1])
    nloc = s1 - s0
    flags = capi.MATH_LITERAL if a.literal else 0

    # a side stream: the legacy default stream's handle is 0, which the C ABI reads as "own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    basin = capi.Basin(P, E, area, nd, region, 1)
    basin.set_stream(stream.cuda_stream)
    par_d = torch.from_numpy(par).to(dev)
    of_d = torch.empty(nloc, dtype=torch.float64, device=dev)
    counts = [a.sets * (r + 1) // world - a.sets * r // world for r in range(world)]
    mx = max(counts)
    of_pad = torch.zeros(mx, dtype=torch.float64, device=dev)       # ragged slices padded for the all-gather
    of_all = torch.empty(mx * world, dtype=torch.float64, device=dev) if world > 1 else of_d

    def step():
        basin.objective_batch_dev(par_d.data_ptr(), nloc, qobs, 36, capi.OF_KGE, flags, of_d.data_ptr())
        if world > 1:
            of_pad[:nloc].copy_(of_d)
            dist.all_gather_into_tensor(of_all, of_pad)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    kern_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    launches0 = capi.kernel_launches()
    e0.record()
    for _ in range(a.steps):
        step()
        kern_ms.append(None)
    e1.record()
    fence()
    launches = capi.kernel_launches() - launches0         # counted by the library at every launch site
    ms_total = e0.elapsed_time(e1)
    last_kernel_ms = basin.last_kernel_ms()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total, last_kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    units_step = float(a.cells) * a.sets * a.months
    value = units_step * a.steps / (ms_total * 1e-3)
    of_host = of_all.cpu().numpy()
    if world > 1:
        of_host = np.concatenate([of_host.reshape(world, mx)[r, :c] for r, c in enumerate(counts)])

    # ---- roofline of the recursion kernel ------------------------------------------------------
    peak_tf = capi.fp64_peak_tflops(4096)
    units_kernel = float(a.cells) * nloc * a.months
    rate = units_kernel / (kernel_ms * 1e-3)
    ach_tf = rate * FLOP_PER_UNIT_ALGO / 1e12                       # algorithmic flops per unit x units / time
    roofline = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "traffic": None, "kernel": "k_gr2m_calib4<cells per thread 2, fixed-point rounding>" if not a.literal else "k_gr2m_calib2<literal>",
                "kernel_ms": kernel_ms, "units_per_launch": units_kernel,
                "algorithmic_flop_per_unit": FLOP_PER_UNIT_ALGO,
                "achieved_def": "SURVEY 8(d) algorithmic count (67 FP64 operations per subbasin x set x month, every add / "
                                "mul / div / compare / exp / cbrt call as 1) x units per launch / CUDA-event kernel time",
                "peak_def": "DFMA micro-benchmark measured in this run (gr2m_fp64_peak: independent chains, 2 flop per DFMA); "
                            "MEASURED_PEAKS.json has no FP64 entry and B200_PROFILING.md states none",
                "pipe_model": pipe_model(units_kernel, kernel_ms, clocks)}
    if NCU and not a.literal:
        # executed (not algorithmic) work from the ncu op counters of the shipped kernel; DRAM bytes scale with the units
        roofline.update({
            "traffic": (NCU.get("full_launch") or NCU)["dram_bytes_per_unit"] * units_kernel,
            "traffic_def": "dram__bytes_read.sum + dram__bytes_write.sum per launch, measured by ncu on the "
                           f"{(NCU.get('full_launch') or NCU)['units_per_launch']:.3g}-unit launch of this kernel "
                           f"({(NCU.get('full_launch') or NCU)['source']}) and scaled by units when the workload differs; "
                           "algorithmic bytes are ~0 (forcing shared by 8 sets per block and by L2, one outlet sum per set and month)",
            "executed_flop_per_unit": NCU["executed_flop_per_unit"],
            "achieved_executed_tflops": rate * NCU["executed_flop_per_unit"] / 1e12,
            "achieved_executed_def": "2 x DFMA + DMUL + DADD thread instructions per unit from the ncu op counters x units / time",
            "fp64_pipe_inst_per_unit": NCU["fp64_pipe_inst_per_unit"], "all_inst_per_unit": NCU["all_thread_inst_per_unit"],
            "fp64_pipe_busy_ncu": NCU["fp64_pipe_busy"], "issue_active_ncu": NCU["issue_active"],
            "smem_pipe_busy_ncu": NCU.get("smem_pipe_busy"),
            "smem_wavefronts_per_warp_unit": NCU.get("smem_wavefronts_per_warp_unit"),
            "limiter_note": "two co-limiters (ncu, profiles/r2_calib4_ncu_full.txt): issue slots (an FP64-pipe instruction takes 2, "
                            "62 % busy) and the shared-memory data pipe (25 wavefronts per warp and unit, 9 of them bank "
                            "conflicts of the random table lookups, 75 % busy)"})

    # ---- end to end through the host-buffer C ABI ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        Pp = torch.from_numpy(P).pin_memory()
        Ep = torch.from_numpy(E).pin_memory()
        parp = torch.from_numpy(par).pin_memory()
        basin.close()
        n_e2e = max(1, min(a.steps, 2))

        e2e_parts = {"create_upload_tile_s": 0.0, "objective_s": 0.0, "destroy_s": 0.0}

        def e2e_step():
            t0 = time.perf_counter()
            if world > 1:
                # the table is replicated: every rank uploads the rows of cells/world subbasins from pinned
                # memory and the ranks all-gather them over NVLink (gr2msemidistr_b200.dist.replicate_forcing)
                Pd, Ed = gdist.replicate_forcing(Pp, Ep, dev)
                b = capi.Basin(None, None, area, nd, region, 1, dev_ptrs=(Pd.data_ptr(), Ed.data_ptr()),
                               shape=(a.cells, a.months), stream=torch.cuda.current_stream().cuda_stream)
                del Pd, Ed
            else:
                b = capi.Basin(Pp.numpy(), Ep.numpy(), area, nd, region, 1)
            t1 = time.perf_counter()
            r = b.objective_batch(parp.numpy(), qobs, 36, capi.OF_KGE, flags, want_crit=False)
            t2 = time.perf_counter()
            b.close()
            t3 = time.perf_counter()
            e2e_parts["create_upload_tile_s"] += t1 - t0
            e2e_parts["objective_s"] += t2 - t1
            e2e_parts["destroy_s"] += t3 - t2
            return r["of"]

        def timed_e2e(pinned):
            e2e_step()
            fence()
            for k_ in e2e_parts:
                e2e_parts[k_] = 0.0
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                of_ = e2e_step()
            torch.cuda.synchronize()
            dt_ = time.perf_counter() - t0
            tt = torch.tensor([dt_], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), of_, {k_: v_ / n_e2e for k_, v_ in e2e_parts.items()}

        dt, of_e2e, parts = timed_e2e(True)
        same = bool(np.array_equal(of_e2e, of_host[s0:s1], equal_nan=True))
        e2e = {"value": units_step * n_e2e / dt, "unit": UNIT,
               "h2d_bytes_per_step": int((P.nbytes + E.nbytes) // world + area.nbytes + nd.nbytes + region.nbytes + par.nbytes
                                         + qobs.nbytes),
               "h2d_note": "per rank" + ("; forcing rows uploaded 1/%d per rank and all-gathered (NCCL)" % world if world > 1 else ""),
               "d2h_bytes_per_step": int(nloc * 8), "steps": n_e2e, "matches_resident_run": same,
               "seconds_per_step": parts,
               "api": ("dist.replicate_forcing + gr2m_basin_create_dev" if world > 1 else "gr2m_basin_create")
                      + " + gr2m_objective_batch + gr2m_basin_destroy (pinned host buffers)"}
        if world == 1:
            # what an R caller has: ordinary pageable matrices (the library stages them through its pinned ring)
            Pp, Ep, parp = torch.from_numpy(P), torch.from_numpy(E), torch.from_numpy(par)
            dtp, of_p, parts_p = timed_e2e(False)
            e2e["pageable"] = {"value": units_step * n_e2e / dtp, "unit": UNIT, "seconds_per_step": parts_p,
                               "matches_resident_run": bool(np.array_equal(of_p, of_host[s0:s1], equal_nan=True)),
                               "note": "same call sequence on pageable host arrays, as .Call from R passes them"}
    else:
        basin.close()

    # ---- simulation mode (Run_GR2MSemiDistr: one parameter set, every series written) -------------
    simulation = None
    if rank == 0:
        try:
            with capi.Basin(P, E, area, nd, region, 1) as bs:
                bs.run(par_all[0], outputs=("QS",))
                t0 = time.perf_counter()
                bs.run(par_all[0])
                dt = time.perf_counter() - t0
                kms = bs.last_kernel_ms()
            u = float(a.cells) * a.months
            simulation = {"metric": "gr2m_cell_months_per_sec (simulation mode, 1 set, PR/AE/SM/RU/QS written)",
                          "workload": f"{a.cells} subbasins x {a.months} months x 1 parameter set",
                          "kernel_ms": kms, "kernel_value": u / (kms * 1e-3),
                          "e2e_value": u / dt, "e2e_note": "gr2m_run on pageable host arrays: kernel writes R's layout, 5 x D2H of "
                                                           "[ntime x ncell] doubles staged through the pinned ring",
                          "roofline": {"bound": "hbm", "achieved": u * 56.0 / (kms * 1e-3) / 1e9, "unit": "GB/s",
                                       "algorithmic_bytes_per_unit": 56,
                                       "note": "launch-latency sized (one thread per subbasin): reported, not tuned"}}
        except Exception as ex:
            simulation = {"error": repr(ex)}

    routed_cal = None
    if rank == 0 and not a.no_routed:
        try:
            routed_cal = bench_routed_calibration(a, capi, torch)
        except Exception as ex:   # secondary figure: report, do not lose the headline line
            routed_cal = {"error": repr(ex)}

    routing, routing_filled = None, None
    if a.route_n > 0 and a.route_terrain in ("tilted", "both"):
        try:
            routing = bench_routing(a, capi, torch, rank, world, dist, "tilted")
        except Exception as ex:   # report, do not lose the headline line
            routing = {"error": repr(ex)}
    if a.route_n > 0 and a.route_terrain in ("filled", "both"):
        try:
            capi.release_cached_memory()
            routing_filled = bench_routing(a, capi, torch, rank, world, dist, "filled")
        except Exception as ex:
            routing_filled = {"error": repr(ex)}

    cpu = None
    if rank == 0 and not a.no_cpu:
        cpu, _, _ = cpu_port(a, os.cpu_count() or 1, a.cpu_seconds)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "sets_per_gpu": nloc, "objective": "KGE", "warmup_months": 36,
                           "l2": "forcing (768 MB) and outlet partials larger than L2; no flush needed",
                           "parallelism": f"parameter sets sharded over {world} GPU(s), forcing replicated",
                           "reference_arm_note": "bench.py --impl reference is the CPU oracle PORT (R / airGR / TauDEM cannot run in "
                                                 "this image) on a bounded DOWN-SCALED sample of this workload (2 000 subbasins x "
                                                 "as many sets as fit ~15 s x all months), rate-extrapolated: a lower bound on the "
                                                 "reference's time, not the same configuration"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches,
                "gpu_launches_note": "kernels of libgr2m_b200 launched by this rank inside the timed region, counted by the library "
                                     "itself (gr2m_kernel_launches); per step: k_gr2m_calib4 + k_objective", "clocks": clocks, "routing": routing, "routing_filled_terrain": routing_filled, "routed_calibration": routed_cal, "simulation": simulation,
                "objective_checksum": float(np.nansum(of_host))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
