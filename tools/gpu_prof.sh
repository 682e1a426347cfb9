# ncu captures of the hot kernels on the C3 stand-in at 256x256 (bounce 1 launches); run after the plain command exited 0
set -x
CMD="python bench.py --steps 1 --warmup 1 --workload c3 --size 256 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_shade -s 1 -c 1 -f -o gpurun_out/prof_shade_r1e $CMD > gpurun_out/ncu_shade.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 1 -c 1 -f -o gpurun_out/prof_closest_r1e $CMD > gpurun_out/ncu_closest.log 2>&1
