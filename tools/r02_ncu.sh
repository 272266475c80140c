#!/bin/bash
# ncu evidence at the shipped revision: (1) launch list of one bench step, (2) --set full of the tensor-bound kernels,
# (3) --set full of one eager decode step (stream GEMMs incl. the folded-norm forms + decode attention)
mkdir -p gpurun_out
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__cluster_size|launch__shared_mem_per_block_dynamic|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --ncu-range --no-gpu-reference > gpurun_out/r02_ncu_range.log 2>&1
tail -1 gpurun_out/r02_ncu_range.log
python tools/ncu_summarize.py gpurun_out/r02_launches.csv --decode-replays 199 --decode-launches 147 > gpurun_out/r02_ncu_launches_bench_step.txt 2>&1
head -30 gpurun_out/r02_ncu_launches_bench_step.txt
gzip -f gpurun_out/r02_launches.csv
summ() {  # rep -> text
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
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_tensor_full python tools/ncu_tensor_kernels.py > /dev/null 2>&1
summ gpurun_out/r02_tensor_full.ncu-rep > gpurun_out/r02_ncu_full_tensor_kernels.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_decode_full -c 160 python tools/decode_ncu.py 2 stream 1 > /dev/null 2>&1
summ gpurun_out/r02_decode_full.ncu-rep > gpurun_out/r02_ncu_full_decode_step.txt
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/r02_decode_full.ncu-rep
wc -l gpurun_out/r02_ncu_full_*.txt; head -40 gpurun_out/r02_ncu_full_tensor_kernels.txt
