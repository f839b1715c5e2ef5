"""GPU box tool: a two-pass batch (tritangents, all paths) through ONE host call on all visible devices vs one device:
results must be bit-identical (the hand-over criteria depend on the path alone), every device runs its own second pass.
Usage: python tests/tools/gpu_multi_two_pass.py [limit]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import hcb200  # noqa: E402,F401
from hcb200 import lib, workloads  # noqa: E402


def main():
    limit = int(sys.argv[1]) if len(sys.argv) > 1 else None
    ndev = torch.cuda.device_count()
    w = workloads.tritangents_total_degree(limit)
    out = []
    for devs in ([0], list(range(ndev))):
        api = lib.load(devices=devs)
        h = w.build(api)
        w.track(api, h)   # warm-up (module load per device)
        r = w.track(api, h)
        tm = lib.timing()
        print(f"{len(devs)} device(s): {w.N} paths, kernel {tm.kernel_ms:.0f} ms (max over devices) = {w.N / tm.kernel_ms * 1e3:,.0f} paths/s, "
              f"second pass {tm.handoff_paths} paths, {tm.handoff_ms:.0f} ms (max), codes {np.bincount(r.return_code).tolist()}", flush=True)
        out.append(r)
        del h
    same = all(np.array_equal(a, b, equal_nan=a.dtype.kind in "fc") for a, b in zip(out[0].arrays(), out[1].arrays()))
    print("all devices vs one device bit-identical:", same)
    assert same


if __name__ == "__main__":
    main()
