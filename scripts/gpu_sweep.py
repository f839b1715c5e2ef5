import sys, os, subprocess, json
cfgs = []
for block, bps in ((64, 4), (64, 8), (32, 16), (128, 2), (128, 4)):
    for rep in (148, 592, 1184):
        cfgs.append((block, bps, rep))
for block, bps, rep in cfgs:
    env = dict(os.environ, HC_B200_BLOCK=str(block), HC_B200_BLOCKS_PER_SM=str(bps))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "1", "--replicas", str(rep), "--no-cpu-baseline"] + sys.argv[1:],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"block {block} blocks/SM {bps} replicas {rep}: value {j['value']:.0f} paths/s e2e {j['e2e']['value']:.0f} ms/step {j['ms_per_step']:.1f} grid {j['config']['grid']} frac {j['roofline']['frac']:.4f}", flush=True)
    except Exception as e:
        print("failed", block, bps, rep, out.stderr[-500:], flush=True)
