CMD="python bench.py --steps 1 --warmup 1 --workload c2 --size 512 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_shade -s 2 -c 1 -f -o gpurun_out/prof_shade_c2 $CMD > gpurun_out/ncu_shade_c2.log 2>&1
