# round 2, GPU call 4 (2 GPUs): feature tests incl. motion blur, the group device (peer assembly), torchrun bench at N=2, group bench
set -x
python -m pytest tests/test_gpu_features.py tests/test_gpu_group.py tests/test_frontend.py tests/test_gpu_cubemap.py -m gpu -q > gpurun_out/r2d_tests.log 2>&1; tail -30 gpurun_out/r2d_tests.log
for WL in c4 c2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 --workload $WL > gpurun_out/r2d_bench_${WL}_n2.json 2> gpurun_out/r2d_bench_${WL}_n2.err; tail -c 1800 gpurun_out/r2d_bench_${WL}_n2.json; tail -3 gpurun_out/r2d_bench_${WL}_n2.err
done
python bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench_c4_group2.json 2> gpurun_out/r2d_bench_c4_group2.err; tail -c 1500 gpurun_out/r2d_bench_c4_group2.json; tail -3 gpurun_out/r2d_bench_c4_group2.err
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench_c4_n1.json 2> gpurun_out/r2d_bench_c4_n1.err; tail -c 600 gpurun_out/r2d_bench_c4_n1.json
