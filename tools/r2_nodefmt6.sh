# round 2, traversal loop experiment 6 (1 GPU): results of finished rays written at the next refill (default) against at retirement (prev = the previous commit's library)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf6_tests.log 2>&1; tail -3 gpurun_out/nf6_tests.log
python tools/ab.py c4 2048 1 default prev default prev 2>&1 | tee gpurun_out/nf6_ab_c4.txt
python tools/ab.py c2 1024 1 default prev 2>&1 | tee gpurun_out/nf6_ab_c2.txt
python tools/c5_ab.py 1e7 24 default prev 2>&1 | tee gpurun_out/nf6_ab_c5.txt
