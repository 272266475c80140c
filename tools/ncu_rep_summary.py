"""Text summary of an .ncu-rep (--set full): the roofline-relevant metrics per profiled launch.
    python tools/ncu_rep_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(r"dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"lts__t_sector_hit_rate\.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__cluster_size|"
                 r"launch__shared_mem_per_block_dynamic|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
cols = [i for i, h in enumerate(hdr) if PAT.fullmatch(h)]
print(f"## {sys.argv[1]}  (ncu --set full --clock-control none; one block per profiled launch)")
for r in rows[2:]:
    print("kernel:", re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r[ki])[:160])
    for i in cols:
        print(f"    {hdr[i]:75s} {r[i]} {units[i]}")
