# round 2, GPU call 3: full GPU test suite, two chunk lanes vs one (bench + A/B of the CTA split)
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; tail -25 gpurun_out/r2c_tests.log
grep -E "^(c2|c3|c4|spheres|atrium):" gpurun_out/r2c_tests.log
python tools/ab.py c4 2048 1 default:lanes=1 default:lanes=2 default:lanes=2,tracectas=10,shadectas=6 default:lanes=2,tracectas=6,shadectas=6 default:lanes=2,tracectas=8,shadectas=4 default:lanes=2,tracectas=12,shadectas=4 2>&1 | tee gpurun_out/r2c_ab_c4.txt
python tools/ab.py c3 1024 1 default:lanes=1 default:lanes=2 default:lanes=2,tracectas=10,shadectas=6 2>&1 | tee gpurun_out/r2c_ab_c3.txt
python tools/ab.py c2 1024 1 default:lanes=1 default:lanes=2 default:lanes=2,tracectas=10,shadectas=6 2>&1 | tee gpurun_out/r2c_ab_c2.txt
python bench.py --steps 3 --warmup 2 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; tail -c 2500 gpurun_out/r2c_bench_c4.json; tail -3 gpurun_out/r2c_bench_c4.err
