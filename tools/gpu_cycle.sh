#!/bin/bash
# one GPU cycle for kernel work: parity subset, bench line, one full ncu capture of the dominant kernel
# usage: tools/gpu_cycle.sh TAG [kernel-regex]
TAG=$1; KRE=${2:-k_patch_ws}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_t.log
tail -n 3 gpurun_out/${TAG}_t.log
timeout 300 python bench.py --no-cpu > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_b.json")); print("ms_step", d["ms_per_step"], "kernel", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_b.err").read()[-2000:])
PY
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 2 -c 1 -f -o gpurun_out/prof_${TAG} python tools/prof_patch.py 1000 4 > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/src_${TAG}.csv 2>/dev/null
rm -f gpurun_out/prof_${TAG}.ncu-rep
