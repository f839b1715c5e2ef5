# 8 B200s of one box: the bench as the driver launches it (weak-scaling headline + strong-scaling C5 sweep)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02b_bench_8gpu.err | tail -1 > gpurun_out/r02b_bench_8gpu.json
tail -3 gpurun_out/r02b_bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_8gpu.json')); print('8 GPUs: value', d['value'], 'e2e', d['e2e']['value'], 'c5', d['scaling_c5'].get('value'), d['scaling_c5'].get('e2e'), d['scaling_c5'].get('error'), d['config']['class_counts'])"
