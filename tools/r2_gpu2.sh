# round 2, GPU call 2: new parity tests (at size, float64 truth, features), ncu launch list + full captures of the three big kernels on C4
set -x
python -m pytest tests/test_gpu_atsize.py tests/test_gpu_features.py tests/test_gpu_cubemap.py -m gpu -q -s 2>&1 | tail -40
B="python bench.py --steps 1 --warmup 1 --no-stats --no-cpu-baseline"
$B > gpurun_out/r2b_plain.json 2> gpurun_out/r2b_plain.err; tail -c 300 gpurun_out/r2b_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_c4.csv $B > gpurun_out/r2b_ncu_list.log 2>&1
for K in k_trace_closest k_trace_shadow k_shade; do
  ncu --set full --import-source on --clock-control none -k regex:$K -s 2 -c 1 -f -o gpurun_out/r2_${K}_c4 $B > gpurun_out/r2b_ncu_$K.log 2>&1; tail -2 gpurun_out/r2b_ncu_$K.log
done
ls -la gpurun_out/*.ncu-rep
