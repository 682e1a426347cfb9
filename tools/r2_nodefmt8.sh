# round 2, traversal loop experiment 8 (1 GPU): the warp's queue slice in shared memory and the duplicate of tbest removed (no spill left in the node phase) against the previous commit's library (prev)
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/nf8_tests.log 2>&1; tail -2 gpurun_out/nf8_tests.log
python tools/ab.py c4 2048 1 default prev default prev 2>&1 | tee gpurun_out/nf8_ab_c4.txt
python tools/c5_ab.py 1e7 24 default prev 2>&1 | tee gpurun_out/nf8_ab_c5.txt
