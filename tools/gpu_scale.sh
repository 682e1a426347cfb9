# usage: bash tools/gpu_scale.sh N  — the driver's launch line for N ranks, C2 (default workload) and the C3 stand-in
N=$1
set -x
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/scale_c2_n$N.json 2> gpurun_out/scale_c2_n$N.err; tail -c 1800 gpurun_out/scale_c2_n$N.json; tail -3 gpurun_out/scale_c2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 12 --warmup 3 --workload c3 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err; tail -c 1800 gpurun_out/scale_c3_n$N.json; tail -3 gpurun_out/scale_c3_n$N.err
