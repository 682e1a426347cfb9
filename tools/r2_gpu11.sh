# round 2, GPU call 11 (2 GPUs): group-device tests, where the group's e2e wall time goes, C5 at N=2
set -x
python -m pytest tests/test_gpu_group.py -m gpu -q 2>&1 | tail -4
python tools/group_e2e_probe.py 2 c4 2>&1 | tail -4
python tools/group_e2e_probe.py 1 c4 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 4 --warmup 2 --workload c5 --tris 10000000 --log2-rays 26 > gpurun_out/r2k_c5_n2.json 2> gpurun_out/r2k_c5_n2.err; tail -c 500 gpurun_out/r2k_c5_n2.json
python bench.py --steps 4 --warmup 2 --workload c5 --tris 10000000 --log2-rays 26 > gpurun_out/r2k_c5_n1.json 2> gpurun_out/r2k_c5_n1.err; tail -c 300 gpurun_out/r2k_c5_n1.json
