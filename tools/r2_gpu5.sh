# round 2, GPU call 5 (8 GPUs): scaling of the cube-map wavefront at N = 8, 4 (C4, C2), C5 soup at 1e7 over 8 ranks, the DLL path and the
# group device's frame assembly on 8 GPUs
set -x
nvidia-smi -L | wc -l
run() { N=$1; WL=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 4 --warmup 2 --workload $WL "$@" > gpurun_out/r2e_${WL}_n$N.json 2> gpurun_out/r2e_${WL}_n$N.err; tail -c 900 gpurun_out/r2e_${WL}_n$N.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2e_${WL}_n$N.err | tail -3; }
run 8 c4
run 4 c4
run 8 c2
run 4 c2
run 8 c3
run 8 c5 --tris 10000000 --log2-rays 26
run 4 c5 --tris 10000000 --log2-rays 26
python tools/group_map_probe.py 8 2048 2>&1 | tail -4
( time bash tools/dll_defaults.sh 8 ) 2>&1 | tail -8
python -m pytest tests/test_gpu_group.py -m gpu -q 2>&1 | tail -5
