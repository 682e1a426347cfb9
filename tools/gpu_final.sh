set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4; do python bench.py --steps 12 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json; done
python bench.py --steps 12 --warmup 3 --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_c2.json 2>&1; tail -c 300 gpurun_out/bench_ref_c2.json
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 14 -c 1 -f -o gpurun_out/prof_closest_c3_1024c python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 12 -c 1 -f -o gpurun_out/prof_closest_c2_1024c python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu-baseline > gpurun_out/ncu_launches_c3.log 2>&1
