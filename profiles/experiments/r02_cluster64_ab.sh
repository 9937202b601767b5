B="timeout -s KILL 300 python bench.py --no-extras --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json")); print("$tag", "%.4e"%d["value"], "frac %.3f"%d["roofline"]["frac"], d["kernel"]["threads"], d["kernel"].get("description","")[:100])
except Exception as e: print("$tag FAILED", e)
PY
}
run r3_c64_double_one --workload double --steps 3 --warmup 2
BR2_TILE_ONE_CTA_THREADS=100 run r3_c64_double_cluster --workload double --steps 3 --warmup 2
run r3_c64_triple_one --workload triple --steps 3 --warmup 2
BR2_TILE_ONE_CTA_THREADS=100 run r3_c64_triple_cluster --workload triple --steps 3 --warmup 2
