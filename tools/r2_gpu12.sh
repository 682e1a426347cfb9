# round 2, GPU call 12 (1 GPU): cache policy of the shading kernel (streams evict-first; shading records evict-last)
set -x
python tools/ab.py c4 2048 1 default noshstream shevict default 2>&1 | tee gpurun_out/r2l_ab_c4.txt
python tools/ab.py c3 1024 1 default noshstream shevict 2>&1 | tee gpurun_out/r2l_ab_c3.txt
python tools/ab.py c2 1024 1 default noshstream shevict 2>&1 | tee gpurun_out/r2l_ab_c2.txt
