"""Summarise an `ncu --set full` report (.ncu-rep) as a markdown table for profiles/:
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [launch indices...]
Reads the raw page with `ncu -i ... --page raw --csv` (run where ncu is installed; no GPU needed)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active, % of elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput, % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU shared-memory wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active, %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, %"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock during capture"),
]


def main():
    rep = sys.argv[1]
    pick = [int(a) for a in sys.argv[2:]]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for n, r in enumerate(body):
        if pick and n not in pick:
            continue
        print(f"### launch {n}: `{r[ki][:90]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {label} (`{m}`) | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main()
