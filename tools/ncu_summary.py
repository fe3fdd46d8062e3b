#!/usr/bin/env python
"""Developer tool: condense an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch [constants.json]] > profiles/<name>.md

With a third argument the per-unit FP64 instruction / flop counts and the DRAM traffic of the FIRST kernel
in the report are also written as JSON (profiles/roofline_constants.json is what bench.py reads).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]
STALLS = "smsp__average_warps_issue_stalled_"


def build_id():
    """rkb_build_id() of the library in the tree: the capture is only meaningful for the build it was taken from"""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        from reak_b200 import _abi
        return _abi.load_library().rkb_build_id().decode()
    except Exception as e:
        return "unknown (%s)" % e


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    json_out = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, unit))
        print("## %s" % d.get("Kernel Name", "?"))
        print("")
        print("source: `%s` (ncu --set full --clock-control none)" % rep)
        print("")
        print("| metric | value | unit |")
        print("|---|---|---|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
        st = sorted(((float(d[k]), k) for k in d if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)
        for v, k in st[:8]:
            print("| stall: %s | %.3f | warps per issue |" % (k[len(STALLS):].replace("_per_issue_active.ratio", ""), v))
        try:
            g = lambda k: float(d[k].replace(",", ""))
            cyc = g("sm__cycles_elapsed.avg")
            fma, mul, add = (g("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o) for o in ("dfma", "dmul", "dadd"))
            peak = g("sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained")
            print("")
            print("derived: FP64 thread-instructions per cycle %.1f of %.0f peak = %.1f %% of the DFMA issue rate"
                  % (fma + mul + add, peak, 100.0 * (fma + mul + add) / peak))
            if units:
                print("derived: per unit (%.0f units/launch): %.0f FP64 instructions, %.0f FP64 flops (DFMA = 2)"
                      % (units, (fma + mul + add) * cyc / units, (2 * fma + mul + add) * cyc / units))
            mb = lambda k: g(k) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u.get(k, "byte"), 1.0)
            traffic = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            print("derived: DRAM traffic %.1f MB per launch" % (traffic / 1e6))
            if json_out and units:
                with open(json_out, "w") as f:
                    json.dump({"kernel": d.get("Kernel Name", "?"), "source": rep, "units_per_launch": units, "build_id": build_id(),
                               "fp64_instr_per_state_step": (fma + mul + add) * cyc / units,
                               "flop_per_state_step": (2 * fma + mul + add) * cyc / units,
                               "dram_traffic_bytes_per_launch": traffic,
                               "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                               "kernel_ms_under_ncu": g("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u.get("gpu__time_duration.sum", "ms"), 1.0)},
                              f, indent=1)
                json_out = None
        except Exception as e:  # metric missing in a reduced set
            print("derived: n/a (%s)" % e)
        print("")


if __name__ == "__main__":
    main()
