"""Summarise .ncu-rep captures into a small markdown table: python tools/ncu_summary.py out.md rep1 rep2 ..."""
import csv, subprocess, sys, io
WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe (DMMA) active % [realtime]", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("DMMA inst % of peak", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    ("FP64 (DFMA) pipe inst % of peak", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("warps active % of max", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("registers/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
]
out = ["| kernel | " + " | ".join(n for n, _ in WANT) + " |", "|---|" + "---|" * len(WANT)]
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    hdr, units = r[0], r[1]
    row = r[2]
    cells = []
    for _, key in WANT:
        if key in hdr:
            i = hdr.index(key)
            cells.append(f"{row[i]} {units[i]}".strip())
        else:
            cells.append("n/a")
    out.append("| " + row[hdr.index("Kernel Name")].split("(")[0] + " | " + " | ".join(cells) + " |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
