# round 2, GPU call 8 (1 GPU): SAH-optimal (DP) BVH8 collapse vs the greedy opening; frames must stay bit-identical
set -x
python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -q -x 2>&1 | tail -4
python tools/ab.py c4 2048 1 default:collapse=0,stats=1 default:collapse=1,ctri=30,stats=1 default:collapse=1,ctri=60,stats=1 default:collapse=1,ctri=100,stats=1 default:collapse=1,ctri=200,stats=1 2>&1 | tee gpurun_out/r2h_ab_stats_c4.txt
python tools/ab.py c4 2048 1 default:collapse=0 default:collapse=1,ctri=30 default:collapse=1,ctri=60 default:collapse=1,ctri=100 default:collapse=1,ctri=150 default:collapse=1,ctri=200 default:collapse=1,ctri=400 default:collapse=1,ctri=100,splitleaves=0 2>&1 | tee gpurun_out/r2h_ab_c4.txt
python tools/ab.py c3 1024 1 default:collapse=0 default:collapse=1,ctri=60 default:collapse=1,ctri=100 default:collapse=1,ctri=200 2>&1 | tee gpurun_out/r2h_ab_c3.txt
python tools/ab.py c2 1024 1 default:collapse=0 default:collapse=1,ctri=60 default:collapse=1,ctri=100 default:collapse=1,ctri=200 2>&1 | tee gpurun_out/r2h_ab_c2.txt
