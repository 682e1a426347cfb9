# round 2, traversal loop experiment 7 (1 GPU): node loads issued before the stack pushes (default), + idle-slot check every second macro step
# (tools/build_variant.py idle2 -DYRT_IDLE_EVERY=2), against the previous commit's library (prev)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf7_tests.log 2>&1; tail -3 gpurun_out/nf7_tests.log
python tools/ab.py c4 2048 1 default prev idle2 default prev idle2 2>&1 | tee gpurun_out/nf7_ab_c4.txt
python tools/ab.py c2 1024 1 default prev idle2 2>&1 | tee gpurun_out/nf7_ab_c2.txt
python tools/c5_ab.py 1e7 24 default prev idle2 2>&1 | tee gpurun_out/nf7_ab_c5.txt
