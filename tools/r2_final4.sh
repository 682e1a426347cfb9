# round 2, 2-GPU run after the node-format change: group-device tests, N = 2 scaling points (torchrun ranks and the in-process group device)
set -x
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2w_group_tests.log 2>&1; tail -2 gpurun_out/r2w_group_tests.log
run() { N=$1; WL=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 4 --warmup 3 --workload $WL --no-cpu-baseline "$@" > gpurun_out/r2w_${WL}_n$N.json 2> gpurun_out/r2w_${WL}_n$N.err; tail -c 200 gpurun_out/r2w_${WL}_n$N.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2w_${WL}_n$N.err | tail -3; }
run 2 c4
run 2 c2
run 2 c5
python bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_c4_group2.json 2> gpurun_out/r2w_c4_group2.err; tail -c 200 gpurun_out/r2w_c4_group2.json
