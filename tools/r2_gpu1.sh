# round 2, GPU call 1: tests, the new default bench (C4 with real textures, cube-map wavefront), per-face A/B, cache-policy A/B, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 2 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; tail -c 3000 gpurun_out/r2a_bench_c4.json; tail -5 gpurun_out/r2a_bench_c4.err
python bench.py --steps 2 --warmup 1 --per-face --no-cpu-baseline > gpurun_out/r2a_bench_c4_perface.json 2> gpurun_out/r2a_bench_c4_perface.err; tail -c 1500 gpurun_out/r2a_bench_c4_perface.json
python bench.py --steps 3 --warmup 2 --workload c3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; tail -c 1500 gpurun_out/r2a_bench_c3.json
python bench.py --steps 3 --warmup 2 --workload c2 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; tail -c 1500 gpurun_out/r2a_bench_c2.json
python tools/ab.py c4 2048 1 default nohint streamonly evictonly 2>&1 | tee gpurun_out/r2a_ab_c4.txt
python tools/ab.py c3 1024 1 default nohint streamonly evictonly 2>&1 | tee gpurun_out/r2a_ab_c3.txt
python tools/ab.py c2 1024 1 default nohint streamonly evictonly 2>&1 | tee gpurun_out/r2a_ab_c2.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref_c4.json 2> gpurun_out/r2a_ref_c4.err; tail -c 800 gpurun_out/r2a_ref_c4.json
