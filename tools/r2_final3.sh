# round 2, final 1-GPU run after the node-format change: full GPU test suite, smoke, bench (C4 default, C3, C2, C1, C5), ncu launch list + full captures of the three kernels
set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r2y_tests.log 2>&1; tail -4 gpurun_out/r2y_tests.log; grep -E "primary rays on|rays vs" gpurun_out/r2y_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2y_bench_c4.json 2> gpurun_out/r2y_bench_c4.err; tail -c 300 gpurun_out/r2y_bench_c4.json; tail -2 gpurun_out/r2y_bench_c4.err
for WL in c3 c2 c1; do python bench.py --steps 5 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/r2y_bench_$WL.json 2> gpurun_out/r2y_bench_$WL.err; tail -c 200 gpurun_out/r2y_bench_$WL.json; done
python bench.py --steps 5 --warmup 3 --workload c5 --no-cpu-baseline > gpurun_out/r2y_bench_c5.json 2> gpurun_out/r2y_bench_c5.err; tail -c 200 gpurun_out/r2y_bench_c5.json
B="python bench.py --steps 1 --warmup 1 --no-stats --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_final_launches_c4.csv $B > gpurun_out/r2y_ncu_list.log 2>&1
for K in k_trace_closest k_trace_shadow k_shade; do
  ncu --set full --import-source on --clock-control none -k regex:$K -s 2 -c 1 -f -o gpurun_out/r2_final_${K}_c4 $B > gpurun_out/r2y_ncu_$K.log 2>&1; tail -1 gpurun_out/r2y_ncu_$K.log
done
