python bench.py --steps 4 --warmup 3 --workload c4 --cpu-sample 256 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
python - <<'PY'
import os, sys, time, subprocess
sys.path.insert(0, os.getcwd())
from tests import dae_scene
d = "/tmp/dae_bench"; dae = dae_scene.write_scene(d, "room", tex_size=256)
for size, spp in ((512, 16), (1024, 64)):
    t = time.time(); r = subprocess.run(["yulio_raytracer_b200/lib/rt_test", dae, str(size), str(spp), "10"], capture_output=True, text=True)
    print(size, spp, "rt_test wall", round(time.time() - t, 3), "s;", r.stdout.strip().splitlines()[-1], r.stderr[-300:])
PY
