"""Summarise an ncu report (ncu -i REP --page raw --csv) into the JSON kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/OUT.json "command line" "note" """
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def main():
    rep, out, cmd, note = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")], "Block Size": r[hdr.index("Block Size")], "Grid Size": r[hdr.index("Grid Size")]}
        for k in KEYS:
            if k in hdr:
                d[k] = (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip()
        st = {}
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.04:
                    st[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 2)
        d["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
        kernels.append(d)
    json.dump({"command": cmd, "note": note, "kernels": kernels}, open(out, "w"), indent=1)
    print(out, len(kernels), "kernels")

if __name__ == "__main__":
    main()
