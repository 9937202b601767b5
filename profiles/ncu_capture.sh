#!/bin/bash
# usage (on the GPU box, via gpurun): profiles/ncu_capture.sh <tag> <kernel-regex> [extra bench args]
# One `ncu --set full` capture of one launch of the step kernel (100 integration steps, 4096 instances).
tag=$1; regex=$2; shift 2
python bench.py --steps 1 --warmup 1 --substeps 100 --no-cpu-baseline "$@" > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -f -o gpurun_out/$tag \
    python bench.py --steps 1 --warmup 1 --substeps 100 --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
