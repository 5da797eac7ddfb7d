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
    json_out = None
    args = sys.argv[2:]
    if "--json" in args:
        i = args.index("--json")
        json_out = args[i + 1]
        args = args[:i] + args[i + 2:]
    for h in args:
        print("# " + h)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    summary = []
    for r in rows[2:]:
        print()
        print("kernel: " + r[ki])
        rec = {"kernel": r[ki]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:85s} {units[i]:>16s} {r[i]}")
                rec[w] = {"value": r[i], "unit": units[i]}
        summary.append(rec)
    if json_out:
        import json

        def to_bytes(rec, key):
            v, u = float(rec[key]["value"].replace(",", "")), rec[key]["unit"].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[u]

        gemms = [x for x in summary if "gemm_kernel" in x["kernel"]]
        per = [{"kernel": x["kernel"], "dram_bytes": to_bytes(x, "dram__bytes_read.sum") + to_bytes(x, "dram__bytes_write.sum"),
                "tensor_pipe_pct": float(x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]["value"])} for x in summary]
        out = {"source": rep, "kernels": per,
               "gemm_dram_bytes_per_launch_mean": sum(to_bytes(x, "dram__bytes_read.sum") + to_bytes(x, "dram__bytes_write.sum") for x in gemms) / max(1, len(gemms))}
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
