# round 2, node-test experiment 4 (1 GPU): byte extraction split between PRMT (alu pipe) and IDP.4A (default: the three far planes; dp00 none, dp0a two, dp15 the near planes, dp3f all)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nf4_tests.log 2>&1; tail -3 gpurun_out/nf4_tests.log
python tools/ab.py c4 2048 1 default dp00 dp0a dp15 dp3f default dp00 2>&1 | tee gpurun_out/nf4_ab_c4.txt
python tools/ab.py c2 1024 1 default dp00 dp3f 2>&1 | tee gpurun_out/nf4_ab_c2.txt
python tools/c5_ab.py 1e7 24 default dp00 dp0a dp3f 2>&1 | tee gpurun_out/nf4_ab_c5.txt
