"""Summarise an `ncu --set full` report (.ncu-rep) into the per-kernel text block kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line 1" "header line 2" ... > profiles/rNN_....txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max.per_second", "launch__grid_size", "launch__cluster_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    for h in sys.argv[2:]:
        print("# " + h)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print()
        print("kernel: " + r[ki])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:85s} {units[i]:>16s} {r[i]}")


if __name__ == "__main__":
    main()
