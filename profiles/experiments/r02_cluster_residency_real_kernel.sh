export BR2_TILE_CHUNKS=1
B="timeout -s KILL 300 python bench.py --no-extras --no-cpu-baseline --workload triple200 --substeps 500 --steps 2 --warmup 2"
for b in 45 46 47 48 49 50; do
  $B --batch $b > gpurun_out/r3_res_$b.json 2> gpurun_out/r3_res_$b.err
  python -c "
import json; d=json.load(open('gpurun_out/r3_res_$b.json')); print($b, d['kernel']['launch_ms'])"
done
