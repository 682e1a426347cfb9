# round 2, final 2-GPU run: N = 2 scaling points of the final code, start-up probe of the group device
set -x
run() { N=$1; WL=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 4 --warmup 2 --workload $WL "$@" > gpurun_out/r2x_${WL}_n$N.json 2> gpurun_out/r2x_${WL}_n$N.err; tail -c 200 gpurun_out/r2x_${WL}_n$N.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2x_${WL}_n$N.err | tail -3; }
run 2 c4
run 2 c2
python bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2x_c4_group2.json 2> gpurun_out/r2x_c4_group2.err; tail -c 200 gpurun_out/r2x_c4_group2.json
python tools/startup_probe.py 2>&1 | head -4
YRT_CFG=gpus=2 python tools/startup_probe.py 2>&1 | head -4
python - <<'PY'
import ctypes, time, threading
rt = ctypes.CDLL("libcudart.so.12")
def ctx(i):
    t = time.perf_counter(); rt.cudaSetDevice(i); rt.cudaFree(None); print(f"context on GPU {i}: {time.perf_counter() - t:.3f} s", flush=True)
t0 = time.perf_counter()
th = [threading.Thread(target=ctx, args=(i,)) for i in range(2)]
[t.start() for t in th]; [t.join() for t in th]
print(f"two contexts in parallel threads: {time.perf_counter() - t0:.3f} s")
PY
