# round 2, last 1-GPU call: any-hit kernel bounded to 8 CTAs/SM (64 registers) against the unbounded build (72 registers, prev); its ncu capture; final bench line
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; tail -2 gpurun_out/r2v_tests.log
python tools/ab.py c4 2048 1 default prev default prev 2>&1 | tee gpurun_out/r2v_ab_c4.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_c4.json 2> gpurun_out/r2v_bench_c4.err; tail -c 200 gpurun_out/r2v_bench_c4.json
B="python bench.py --steps 1 --warmup 1 --no-stats --no-cpu-baseline"
ncu --set full --import-source on --clock-control none -k regex:k_trace_shadow -s 2 -c 1 -f -o gpurun_out/r2_final_k_trace_shadow_c4 $B > gpurun_out/r2v_ncu_shadow.log 2>&1; tail -1 gpurun_out/r2v_ncu_shadow.log
