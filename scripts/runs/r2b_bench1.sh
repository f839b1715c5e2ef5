mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -c 600 gpurun_out/r2b_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_default.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"), "frac", d["roofline"]["frac"], "h2d", d["e2e"]["h2d_bytes_per_step"], "launches", d["gpu_launches"], d["e2e"]["entry_point"], d["e2e"]["results_identical_to_resident_arm"])
for k,v in d.get("per_config",{}).items():
    print(k, {kk: v.get(kk) for kk in ("value","e2e","e2e_over_cpu","roofline_frac","second_pass","error")}, (v.get("cpu") or {}).get("value"))
print(d.get("scaling_c5"))
PY
