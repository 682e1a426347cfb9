# round 2, GPU call 6 (1 GPU): full GPU test suite (at-size parity numbers printed), C5 at N=1 (1e7, 1e8), ncu of the C5 traversal launch
set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r2f_tests.log 2>&1; tail -6 gpurun_out/r2f_tests.log
grep -E "primary rays on|rays vs" gpurun_out/r2f_tests.log
python bench.py --workload c5 --tris 10000000 --log2-rays 26 --steps 4 --warmup 2 > gpurun_out/r2f_c5_1e7_n1.json 2> gpurun_out/r2f_c5_1e7_n1.err; tail -c 700 gpurun_out/r2f_c5_1e7_n1.json; tail -2 gpurun_out/r2f_c5_1e7_n1.err
python bench.py --workload c5 --tris 100000000 --log2-rays 26 --steps 4 --warmup 2 > gpurun_out/r2f_c5_1e8_n1.json 2> gpurun_out/r2f_c5_1e8_n1.err; tail -c 700 gpurun_out/r2f_c5_1e8_n1.json; tail -2 gpurun_out/r2f_c5_1e8_n1.err
ncu --set full --import-source on --clock-control none -k regex:k_trace_user -s 3 -c 1 -f -o gpurun_out/r2_k_trace_user_c5_1e7 python bench.py --workload c5 --tris 10000000 --log2-rays 24 --steps 2 --warmup 1 > gpurun_out/r2f_ncu_c5.log 2>&1; tail -2 gpurun_out/r2f_ncu_c5.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; tail -c 400 gpurun_out/r2f_bench_c4.json; tail -3 gpurun_out/r2f_bench_c4.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2f_ref_c4.json 2> gpurun_out/r2f_ref_c4.err; tail -c 300 gpurun_out/r2f_ref_c4.json
python -c "import __graft_entry__ as g; g.smoke()"
