mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "device_side_counts or sweep" 2>&1 | tail -3
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_counts.json 2> gpurun_out/r2b_bench_counts.err; tail -2 gpurun_out/r2b_bench_counts.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_counts.json')); print(d['value'], d['e2e']['value']); print({k:v for k,v in d['scaling_c5'].items() if k!='class_counts'})"
