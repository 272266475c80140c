#!/bin/bash
# final check: whole GPU suite + smoke, decode step, bench (no gpu_reference leg: profiles/r02_bench_1gpu_head.log has it)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02_gpu_tests_final2.log
T1_BENCH_VARIANTS='[{}, {}]' timeout 300 python tools/decode_bench.py 2 200 2>&1 | tail -2 | cut -c1-215 | tee gpurun_out/r02_decode_step_final.txt
timeout 600 python bench.py --steps 5 --warmup 3 --phase-times --no-gpu-reference > gpurun_out/r02_bench_final2.log 2> gpurun_out/r02_bench_final2.err
python - <<'PY'
import json
l = [x for x in open('gpurun_out/r02_bench_final2.log') if x.startswith('{')][-1]
d = json.loads(l)
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline_decode"]["ms"], d["roofline_decode"]["frac"], d["attn"]["ms"], d["attn"]["in_graph"])
PY
tail -2 gpurun_out/r02_bench_final2.err | cut -c1-400
