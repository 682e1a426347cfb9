# round 2, GPU call 10 (1 GPU): config 5 — L2 prefetch of hit children (auto for BVHs larger than L2) and 8 CTAs/SM for the raw traversal kernel
set -x
python tools/c5_ab.py 1e6 24 default:prefetch=0 default:prefetch=1 user8:prefetch=0 user8:prefetch=1 2>&1 | grep -v "^+" | tee gpurun_out/r2j_c5_1e6.txt
python tools/c5_ab.py 1e7 26 default:prefetch=0 default:prefetch=1 user8:prefetch=0 user8:prefetch=1 2>&1 | tee gpurun_out/r2j_c5_1e7.txt
python tools/c5_ab.py 1e8 26 default:prefetch=0 default:prefetch=1 user8:prefetch=1 2>&1 | tee gpurun_out/r2j_c5_1e8.txt
python tools/ab.py c4 2048 1 default default:prefetch=1 2>&1 | tee gpurun_out/r2j_ab_c4.txt
