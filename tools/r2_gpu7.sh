# round 2, GPU call 7 (1 GPU): tuning A/B on the C4 workload with the real textures (the r1 tuning ran on the procedural stand-ins)
set -x
python tools/ab.py c4 2048 1 default default:refill=4 default:refill=6 default:refill=12 default:refill=16 default:trinum=1,triden=1 default:trinum=2,triden=1 default:trinum=4,triden=1 default:trinum=6,triden=1 default:tracectas=7 default:shadectas=5 nosync sync2 shade7:shadectas=7 shade8:shadectas=8 2>&1 | tee gpurun_out/r2g_ab_c4.txt
python tools/ab.py c2 1024 1 default nosync sync2 default:refill=4 default:refill=16 2>&1 | tee gpurun_out/r2g_ab_c2.txt
python tools/ab.py c3 1024 1 default nosync sync2 2>&1 | tee gpurun_out/r2g_ab_c3.txt
