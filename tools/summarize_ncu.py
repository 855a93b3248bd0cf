#!/usr/bin/env python3
"""Turn an .ncu-rep (brought back in gpurun_out/) into a small text summary for profiles/.
  python tools/summarize_ncu.py gpurun_out/r1_k1_match.ncu-rep profiles/r1_k1_match.txt "note"
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = ["# " + rep.split("/")[-1], note, ""]
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    h, u, v = rows[0], rows[1], rows[2]
    lines.append("kernel: " + v[h.index("Kernel Name")] if "Kernel Name" in h else "")
    lines.append("## raw metrics (one launch, ncu --set full --clock-control none)")
    for n in WANT:
        if n in h:
            i = h.index(n)
            lines.append("%-64s %s %s" % (n, v[i], u[i]))
    rd = float(v[h.index("dram__bytes_read.sum")]); wr = float(v[h.index("dram__bytes_write.sum")])
    ur, uw = u[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_write.sum")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    lines.append("dram traffic (read+write): %.3f GB" % ((rd * scale[ur] + wr * scale[uw]) / 1e9))
    lines.append("")
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv"))))
    hdr, data = rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    lines.append("## warp stall samples by reason (all SASS lines)")
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    for s, n in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        lines.append("%-28s %10d  %5.1f%%" % (s, n, 100.0 * n / tot))
    lines.append("")
    lines.append("## top SASS lines by samples: samples, executions, avg active threads, instruction, top stall")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:18]:
        st = max(((int(r[ix[s]] or 0), s) for s in stalls))
        lines.append("%8s %12s %5s  %-62s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]],
                                                   r[ix["Avg. Threads Executed"]], r[ix["Source"]][:62], st[1]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
