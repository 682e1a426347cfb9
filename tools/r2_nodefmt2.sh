# round 2, node-test experiment 2 (1 GPU): hit mask from the sign of tmax - tmin (default) against FSETP/SEL (nf1), without the evict_last policy register (nopol), traversal tuning
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf2_tests.log 2>&1; tail -3 gpurun_out/nf2_tests.log
python tools/ab.py c4 2048 1 default nf1 nopol default:refill=6 default:trinum=2 default:trinum=4 default:tracectas=7 default nf1 2>&1 | tee gpurun_out/nf2_ab_c4.txt
python tools/ab.py c2 1024 1 default nf1 nopol 2>&1 | tee gpurun_out/nf2_ab_c2.txt
python tools/c5_ab.py 1e7 24 default nf1 nopol 2>&1 | tee gpurun_out/nf2_ab_c5.txt
