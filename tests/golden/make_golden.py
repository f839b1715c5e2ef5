"""Extracts the numeric golden vectors of the reference's own test cases into JSON fixtures.

Run in the build container (where /root/reference exists); the GPU box only reads the JSON.
Sources: test/test_cases/steiner_higher_prec.jl (p :5-36, q :38-69, s_p :71-88, s_q :89-105),
         test/test_cases/four_bar.jl (s :27-52, p :53-?, q ..:88)."""
import json
import os
import re

REF = "/root/reference/test/test_cases"
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"


def arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"(\w+)\s*=\s*(?:Complex\{Float64\})?\[\n(.*?)\n\s*\]", txt, re.S):
        vals = []
        for item in m.group(2).split(","):
            item = item.strip()
            if not item:
                continue
            c = re.fullmatch(rf"({NUM})\s*([-+])\s*(\d+\.?\d*(?:[eE][-+]?\d+)?)im", item)
            if c:
                re_, sg, im_ = float(c.group(1)), c.group(2), float(c.group(3))
                vals.append([re_, im_ if sg == "+" else -im_])
            else:
                vals.append([float(item), 0.0])
        out[m.group(1)] = vals
    return out


if __name__ == "__main__":
    st = arrays(os.path.join(REF, "steiner_higher_prec.jl"))
    assert [len(st[k]) for k in ("p", "q", "s_p", "s_q")] == [30, 30, 15, 15], {k: len(v) for k, v in st.items()}
    json.dump(st, open(os.path.join(HERE, "steiner_higher_prec.json"), "w"), indent=0)
    fb = arrays(os.path.join(REF, "four_bar.jl"))
    assert [len(fb[k]) for k in ("s", "p", "q")] == [24, 16, 16], {k: len(v) for k, v in fb.items()}
    json.dump(fb, open(os.path.join(HERE, "four_bar.json"), "w"), indent=0)
    print("ok")
