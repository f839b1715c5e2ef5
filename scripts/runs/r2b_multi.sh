mkdir -p gpurun_out
{
timeout 900 python tests/tools/gpu_multi_two_pass.py
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "all_devices or two_pass" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_2gpu.json')); print('2 GPUs: value', d['value'], 'e2e', d['e2e']['value'], 'c5', d['scaling_c5']['value'], d['scaling_c5']['e2e'])"
} 2>&1 | tee gpurun_out/r02b_multi_gpu.txt
