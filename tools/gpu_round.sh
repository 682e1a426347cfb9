set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3500 gpurun_out/bench_c2.json
python bench.py --steps 12 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3500 gpurun_out/bench_c3.json
python bench.py --steps 4 --warmup 3 --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 1500 gpurun_out/bench_c1.json
# launch list + one full capture of the dominant kernel at the benchmark configuration (after the plain runs above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 12 -c 1 -f -o gpurun_out/prof_closest_c2_1024 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 14 -c 1 -f -o gpurun_out/prof_closest_c3_1024 python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
