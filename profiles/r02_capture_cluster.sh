T="timeout -s KILL 500"
B="python bench.py --no-extras --no-cpu-baseline"
tag=r02_ncu_tile_triple200_cluster
$B --steps 1 --warmup 1 --substeps 100 --workload triple200 --batch 1024 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
$T ncu --set full --clock-control none --import-source on -k regex:br2_step_tile_kernel -s 2 -c 1 -f -o gpurun_out/$tag $B --steps 1 --warmup 1 --substeps 100 --workload triple200 --batch 1024 > gpurun_out/${tag}_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/$tag.ncu-rep 3072000 > gpurun_out/$tag.txt 2>&1
rm -f gpurun_out/$tag.ncu-rep
