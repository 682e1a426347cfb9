# round 2, node-test experiment 3 (1 GPU): node/triangle phase vote threshold after the cheaper node test; pipe microbenchmark
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_pipes tools/mb/pipes.cu && /tmp/mb_pipes 2>&1 | tee gpurun_out/nf3_mb_pipes.txt
python tools/ab.py c4 2048 1 default default:trinum=2 default:trinum=1 default:trinum=3:triden=2 default:trinum=5:triden=2 default:trinum=2:refill=6 default:trinum=2:refill=12 2>&1 | tee gpurun_out/nf3_ab_c4.txt
python tools/ab.py c3 1024 1 default default:trinum=2 default:trinum=1 default:trinum=3:triden=2 2>&1 | tee gpurun_out/nf3_ab_c3.txt
python tools/ab.py c2 1024 1 default default:trinum=2 default:trinum=1 default:trinum=3:triden=2 2>&1 | tee gpurun_out/nf3_ab_c2.txt
python tools/c5_ab.py 1e7 24 default default:trinum=2 default:trinum=1 default:trinum=3:triden=2 2>&1 | tee gpurun_out/nf3_ab_c5.txt
