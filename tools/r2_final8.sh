# round 2, final 8-GPU run: scaling of the final code at N = 8 and 4 (C4, C2), N = 8 (C3, C5 1e7), and the DLL path on 8 GPUs
set -x
run() { N=$1; WL=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 4 --warmup 2 --workload $WL "$@" > gpurun_out/r2y_${WL}_n$N.json 2> gpurun_out/r2y_${WL}_n$N.err; tail -c 300 gpurun_out/r2y_${WL}_n$N.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2y_${WL}_n$N.err | tail -3; }
run 8 c4
run 4 c4
run 8 c2
run 4 c2
run 8 c3
run 8 c5 --tris 10000000 --log2-rays 26
run 4 c5 --tris 10000000 --log2-rays 26
( time bash tools/dll_defaults.sh 8 ) 2>&1 | tail -6
