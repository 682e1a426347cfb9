set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.json
python bench.py --steps 12 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 400 gpurun_out/bench_c3.json
python bench.py --steps 12 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 400 gpurun_out/bench_c4.json
python bench.py --steps 12 --warmup 3 --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 14 -c 1 -f -o gpurun_out/prof_closest_c3_1024b python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_shade -s 14 -c 1 -f -o gpurun_out/prof_shade_c3_1024b python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_c3s.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_shade -s 12 -c 1 -f -o gpurun_out/prof_shade_c2_1024b python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2s.log 2>&1
