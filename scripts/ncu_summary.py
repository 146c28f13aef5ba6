"""Summarise an .ncu-rep (read here, no GPU needed):
    python scripts/ncu_summary.py rep.ncu-rep units_per_launch [counters.json [full_traffic.csv full_units]] > profiles/x.txt
With a third argument the per-unit counters bench.py quotes (instructions, executed flops, DRAM bytes, pipe
utilisation) are also written as JSON; a fourth and fifth add the DRAM traffic of the full-size launch from the
four-metric ncu pass (--csv --log-file) as "full_launch"."""
import csv, io, json, os, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None
jout = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, unit))
    print("=" * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:80s} {d[k]:>22s} {u.get(k, '')}")
    print("-- stall reasons (warps per issue-active cycle) --")
    for k in sorted(d):
        if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"   {k[len(STALL):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
    if units:
        cyc = float(d["sm__cycles_elapsed.max"].replace(",", ""))
        f = sum(float(d[f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed"].replace(",", "")) for o in ("dfma", "dmul", "dadd"))
        tot = float(d["smsp__inst_executed.sum"].replace(",", "")) * float(d["smsp__thread_inst_executed_per_inst_executed.ratio"])
        print(f"-- derived for {units:.4g} units per launch --")
        print(f"   DFMA+DMUL+DADD thread instructions per unit : {f * cyc / units:8.2f}")
        print(f"   all thread instructions per unit            : {tot / units:8.2f}")
        print(f"   units per second                            : {units / (float(d['gpu__time_duration.sum'].replace(',', '')) * 1e-3):.4g}" if u.get('gpu__time_duration.sum') == 'ms' else "")
        if jout:
            g = lambda k: float(d[k].replace(",", ""))
            per = {o: g(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") * cyc / units for o in ("dfma", "dmul", "dadd")}
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            dram = sum(g(k) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            busy = g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0
            # FP64-pipe warp instructions: the pipe is busy 2 cycles per warp instruction on each of the 4 sub-partitions
            pipe_inst = busy * g("sm__cycles_active.avg") * 4 / 2.0 * 32 * g("launch__sm_count") / units if "sm__cycles_active.avg" in d and "launch__sm_count" in d else None
            full = None
            if len(sys.argv) > 5:
                fr = {r[12]: (float(r[14].replace(",", "")), r[13]) for r in csv.reader(open(sys.argv[4])) if len(r) > 14 and r[0].isdigit()}
                fu = float(sys.argv[5])
                fb = fr["dram__bytes_read.sum"][0] + fr["dram__bytes_write.sum"][0]
                full = {"source": os.path.basename(sys.argv[4]) + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... on the "
                                  f"{fu:.3g}-unit launch of the default bench workload)",
                        "units_per_launch": fu, "dram_bytes_per_launch": fb, "dram_bytes_per_unit": fb / fu,
                        "kernel_ms_under_ncu": fr["gpu__time_duration.sum"][0] * (1e-6 if fr["gpu__time_duration.sum"][1] == "ns" else 1.0),
                        "fp64_pipe_busy": fr["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0] / 100.0}
            json.dump({"full_launch": full, "source": os.path.basename(rep), "kernel": d["Kernel Name"], "units_per_launch": units,
                       "kernel_ms_under_ncu": g("gpu__time_duration.sum"),
                       "dfma_per_unit": per["dfma"], "dmul_per_unit": per["dmul"], "dadd_per_unit": per["dadd"],
                       "executed_flop_per_unit": 2 * per["dfma"] + per["dmul"] + per["dadd"],
                       "fp64_pipe_inst_per_unit": pipe_inst if pipe_inst else f * cyc / units,
                       "fp64_pipe_inst_def": "DFMA + DMUL + DADD + DSETP: FP64-pipe busy cycles / 2 per warp instruction x 32 lanes / units" if pipe_inst else "DFMA + DMUL + DADD only",
                       "all_thread_inst_per_unit": tot / units,
                       "smem_wavefronts_per_warp_unit": (g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") * 32 / units
                                                         if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in d else None),
                       "smem_pipe_busy": (g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed") / 100.0
                                          if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed" in d else None),
                       "fp64_pipe_busy": busy, "issue_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                       "xu_pipe_busy": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active") / 100.0,
                       "dram_bytes_per_launch": dram, "dram_bytes_per_unit": dram / units,
                       "registers": int(g("launch__registers_per_thread")), "smem_per_block_kb": g("launch__shared_mem_per_block_dynamic")},
                      open(jout, "w"), indent=1)
