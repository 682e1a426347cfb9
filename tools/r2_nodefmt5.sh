# round 2, traversal loop experiment 5 (1 GPU): idle-slot mask from the vote's ballots (default at the time; reverted) against the previous loop
# (prev = the library of the previous commit copied to lib/variants/libyrt_prev.so); shared-memory stack depth 4 / 8 / 12 (tools/build_variant.py smst4 -DYRT_SM_STACK=4 ...)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf5_tests.log 2>&1; tail -3 gpurun_out/nf5_tests.log
python tools/ab.py c4 2048 1 default prev smst4 smst12 default prev 2>&1 | tee gpurun_out/nf5_ab_c4.txt
python tools/ab.py c2 1024 1 default prev 2>&1 | tee gpurun_out/nf5_ab_c2.txt
python tools/c5_ab.py 1e7 24 default prev smst4 smst12 2>&1 | tee gpurun_out/nf5_ab_c5.txt
