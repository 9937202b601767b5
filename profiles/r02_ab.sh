#!/bin/bash
# usage (on the GPU box): profiles/r02_ab.sh <tag> [lib]  -- bench lines of the multi-segment workloads with one build of the library
tag=$1; lib=$2
[ -n "$lib" ] && export BR2_CUDA_LIB=$PWD/$lib
B="timeout -s KILL 300 python bench.py --no-extras --no-cpu-baseline"
$B --steps 3 --warmup 3 > gpurun_out/${tag}_single.json 2> gpurun_out/${tag}_single.err
$B --workload double --steps 3 --warmup 2 > gpurun_out/${tag}_double.json 2> gpurun_out/${tag}_double.err
$B --workload triple --steps 3 --warmup 2 > gpurun_out/${tag}_triple.json 2> gpurun_out/${tag}_triple.err
$B --workload triple200 --batch 1024 --steps 3 --warmup 2 > gpurun_out/${tag}_t200x1024.json 2> gpurun_out/${tag}_t200x1024.err
$B --workload triple200 --substeps 1000 --steps 3 --warmup 2 > gpurun_out/${tag}_t200x8192.json 2> gpurun_out/${tag}_t200x8192.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.load(open(f))
        print(f, "%.4e" % d["value"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], d["kernel"]["threads"], d["kernel"]["registers"])
    except Exception as exc:
        print(f, "FAILED", exc)
PY
