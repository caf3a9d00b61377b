"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed) into a small markdown table for profiles/."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64_insts"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "barrier", "not_selected", "selected", "dispatch_stall", "branch_resolving", "no_instruction", "sleeping",
          "membar", "tex_throttle", "drain", "misc"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep}`  (ncu --set full --clock-control none; per launch)\n")
    for r in data:
        name = r[idx["Kernel Name"]]
        m = re.search(r"rbgs_stream_kernel<([^>]*)>", name)
        short = f"rbgs_stream_kernel<{m.group(1)}>" if m else name[:80]
        print(f"## {short}   grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in idx:
                print(f"| {label} ({k}) | {r[idx[k]]} | {units[idx[k]]} |")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in idx:
                try:
                    st.append((float(r[idx[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("\nwarp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{s} {v:.2f}" for v, s in st[:7]) + "\n")


if __name__ == "__main__":
    main()
