# 2-GPU box: the whole GPU suite (test_all_devices_through_one_call then really uses two devices) + the 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/gputest_2gpu.log 2>&1; tail -8 gpurun_out/gputest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; tail -c 1500 gpurun_out/bench_r02_2gpu.json; tail -3 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
# one host call over both devices vs one device: throughput of the sharded call
import os, sys, time
sys.path[:0] = [os.getcwd()]
import numpy as np, hcb200
from hcb200 import lib, workloads
for devs in ([0], [0, 1]):
    api = lib.load(devices=devs)
    w = workloads.cyclic_polyhedral(7, 480)
    h = w.build(api)
    w.track(api, h)
    t0 = time.perf_counter(); r = w.track(api, h); dt = time.perf_counter() - t0
    tm = lib.timing()
    print(f"one call, devices {devs}: {w.N / dt:,.0f} paths/s end to end (kernel max {tm.kernel_ms:.0f} ms, {tm.devices} device(s)), success {(r.return_code == 1).sum()}", flush=True)
    del h
PY
