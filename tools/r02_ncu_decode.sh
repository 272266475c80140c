#!/bin/bash
# ncu --set full of one eager decode step at the shipped revision (stream GEMMs with the batched fix-up, folded-norm forms, decode attention)
mkdir -p gpurun_out
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__cluster_size|launch__shared_mem_per_block_dynamic|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum'
summ() {
  ncu -i $1 --page raw --csv 2>/dev/null | python - "$M" <<'PY'
import csv, re, sys
pat = re.compile(sys.argv[1])
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
cols = [i for i, h in enumerate(hdr) if pat.fullmatch(h)]
for r in rows[2:]:
    print("kernel:", r[ki][:150])
    for i in cols:
        print(f"    {hdr[i]:75s} {r[i]} {units[i]}")
PY
}
timeout 500 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_decode_full_v2 -c 40 python tools/decode_ncu.py 2 stream 1 > /dev/null 2>&1
echo "## gpurun_out/r02_decode_full_v2.ncu-rep  (ncu --set full --clock-control none, first 40 launches of one eager decode step at HEAD; one block per profiled launch)" > gpurun_out/r02_ncu_full_decode_step_v2.txt
summ gpurun_out/r02_decode_full_v2.ncu-rep >> gpurun_out/r02_ncu_full_decode_step_v2.txt
rm -f gpurun_out/r02_decode_full_v2.ncu-rep
wc -l gpurun_out/r02_ncu_full_decode_step_v2.txt; grep -A4 "stream_gemm" gpurun_out/r02_ncu_full_decode_step_v2.txt | head -40
