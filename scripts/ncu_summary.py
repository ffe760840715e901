"""Summarise an .ncu-rep (read on the CPU box): key metrics of each profiled kernel + top stall lines.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--source N]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    return hdr, units, vals


def main():
    rep = sys.argv[1]
    hdr, units, vals = raw(rep)
    for v in vals:
        d = dict(zip(hdr, v))
        u = dict(zip(hdr, units))
        print("## %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        print("| metric | unit | value |\n|---|---|---|")
        for m in METRICS:
            if m in d:
                print("| %s | %s | %s |" % (m, u[m], d[m]))
        stalls = [(k, d[k]) for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
        stalls = sorted(((k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(x.replace(",", "") or 0)) for k, x in stalls),
                        key=lambda t: -t[1])
        print("stall samples:", ", ".join("%s %d" % s for s in stalls[:9]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[0]
        try:
            i_samp = h.index("# Samples")
        except ValueError:
            i_samp = [i for i, x in enumerate(h) if "Samples" in x][0]
        i_src = h.index("Source") if "Source" in h else 1
        i_inst = [i for i, x in enumerate(h) if x.strip() == "# Instructions Executed" or "Instructions Executed" in x]
        body = [r for r in rows[1:] if len(r) > i_samp]
        def num(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        body.sort(key=lambda r: -num(r[i_samp]))
        print("\ntop %d source/SASS lines by samples (%s):" % (n, h[i_samp]))
        for r in body[:n]:
            extra = r[i_inst[0]] if i_inst else ""
            print("%8s  %10s  %s" % (r[i_samp], extra, r[i_src][:140]))


if __name__ == "__main__":
    main()
