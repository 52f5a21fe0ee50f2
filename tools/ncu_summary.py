"""Summarise .ncu-rep captures into a small markdown table: python tools/ncu_summary.py out.md rep1 rep2 ...
A capture of hqr_kernel also writes <dir of out.md>/hqr_traffic.json (DRAM bytes per launch: bench.py's roofline.traffic)."""
import csv, json, os, subprocess, sys, io
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
    name = row[hdr.index("Kernel Name")].split("(")[0]
    out.append("| " + name + " | " + " | ".join(cells) + " |")
    if name.startswith("hqr_kernel"):
        def val(key):
            i = hdr.index(key)
            v = float(row[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
        grid = int(float(row[hdr.index("launch__grid_size")].replace(",", "")))
        json.dump({"kernel": "hqr_kernel", "m": 1024, "members_per_launch": grid,
                   "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                   "duration_ns_under_ncu": row[hdr.index("gpu__time_duration.sum")],
                   "source": "ncu --set full capture " + os.path.basename(rep) + " (tools/ncu_pass.sh, bench --members 148, m = 1024)"},
                  open(os.path.join(os.path.dirname(os.path.abspath(sys.argv[1])), "hqr_traffic.json"), "w"), indent=1)
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
