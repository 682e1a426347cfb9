set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_c2_s2.json 2> gpurun_out/bench_c2_s2.err; tail -c 3000 gpurun_out/bench_c2_s2.json
python bench.py --steps 12 --warmup 3 --workload c3 > gpurun_out/bench_c3_s2.json 2> gpurun_out/bench_c3_s2.err; tail -c 3000 gpurun_out/bench_c3_s2.json
