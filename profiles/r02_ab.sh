#!/bin/bash
# usage (on the GPU box): profiles/r02_ab.sh <tag> [lib] [workloads...]  -- bench lines of the workloads with one build of the library
tag=$1; lib=$2; shift 2
[ -n "$lib" ] && [ "$lib" != "-" ] && export BR2_CUDA_LIB=$PWD/$lib
B="timeout -s KILL 300 python bench.py --no-extras --no-cpu-baseline"
W="${@:-single double triple t200x1024 t200x8192}"
for w in $W; do
  case $w in
    single) a="--steps 3 --warmup 3";;
    double) a="--workload double --steps 3 --warmup 2";;
    triple) a="--workload triple --steps 3 --warmup 2";;
    t200x1024) a="--workload triple200 --batch 1024 --steps 3 --warmup 2";;
    t200x8192) a="--workload triple200 --substeps 1000 --steps 3 --warmup 2";;
  esac
  $B $a > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.load(open(f))
        print(f, "%.4e" % d["value"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], d["kernel"]["threads"], d["kernel"]["registers"])
    except Exception as exc:
        print(f, "FAILED", exc)
PY
