# round 2, node-format experiment (1 GPU): new node format (one triMask per node, slot-bit hit mask + permutation table) + packed FP32 node test
# against the previous library (lib/variants/libyrt_old.so) and the new format with the scalar node test (nf0)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf_tests.log 2>&1; tail -5 gpurun_out/nf_tests.log
python tools/ab.py c4 2048 1 default old nf0 default old 2>&1 | tee gpurun_out/nf_ab_c4.txt
python tools/ab.py c3 1024 1 default old nf0 2>&1 | tee gpurun_out/nf_ab_c3.txt
python tools/ab.py c2 1024 1 default old nf0 2>&1 | tee gpurun_out/nf_ab_c2.txt
python tools/c5_ab.py 1e7 24 default old nf0 2>&1 | tee gpurun_out/nf_ab_c5.txt
