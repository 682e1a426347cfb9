set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json
python bench.py --steps 12 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.json
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; tail -c 900 gpurun_out/bench_ref_c2.json
CMD="python bench.py --steps 1 --warmup 1 --workload c3 --size 256 --no-cpu-baseline"
ncu --set full --import-source on --clock-control none -k regex:k_shade -s 1 -c 1 -f -o gpurun_out/prof_shade_r1d $CMD > gpurun_out/ncu_shade.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_shadow -s 1 -c 1 -f -o gpurun_out/prof_shadow_r1d $CMD > gpurun_out/ncu_shadow.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_launches_c3.log 2>&1
