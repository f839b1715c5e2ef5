"""Static instruction footprint of a kernel by source function (no GPU needed):
    cuobjdump -xelf all libhc_b200.so ; nvdisasm -gi -c hc_api.sm_100a.cubin > all.sass
    python scripts/sass_static_by_function.py all.sass hc_track_tpl_kernelILi12288
Counts the SASS instructions (16 bytes each) attributed by -lineinfo to source functions -- where the 1.26 MB of the
thread-per-path kernel come from, which matters because instruction fetch is a visible stall.
An instruction is charged to the innermost frame of its inlining chain that is not an arithmetic / accessor helper
(hc_common.h, hc_coop.h, the slot accessors of hc_tape.h), i.e. to the algorithmic function that contains it."""
import bisect, collections, os, re, sys

sass, kernel = sys.argv[1], sys.argv[2]
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "homotopycontinuation.jl_b200", "csrc")


def marks(path):
    out = []
    for i, l in enumerate(open(path).read().split('\n'), 1):
        m = re.match(r'\s*(template <[^>]*>\s*)?HC_(HDN|HD|D)\s+(static\s+)?[\w:<>&\*,\s]+?\s+(\w+)\(', l)
        if m: out.append((i, m.group(4)))
    return out


HELPERS = {"cx", "int", "double", "get", "adv", "mk", "PCur", "pld", "tload", "tstore", "ser_load", "ser_store", "mneg", "mfma",
           "fop_eval", "t_mul", "t_add", "t_sqr", "LRef", "SV", "operator", "at", "ld_at", "st_at", "store_in"}
mk = {}
inside, cur = False, None
chain, chain_open = [], False


def fn_of(f, ln):
    if f not in mk:
        p = os.path.join(csrc, f)
        mk[f] = marks(p) if os.path.exists(p) else []
    starts = [a for a, _ in mk[f]]
    j = bisect.bisect_right(starts, ln) - 1
    return mk[f][j][1] if j >= 0 else "?"
per_fn, per_line = collections.Counter(), collections.Counter()
total = 0
for l in open(sass, errors="replace"):
    if l.startswith("\t.section"):
        inside = (".text." in l) and (kernel in l)
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', l)
    if m:
        fr = (os.path.basename(m.group(1)), int(m.group(2)))
        if chain and chain_open and chain[-1][1] == fr:   # continuation of the chain: (frame, its call site)
            chain.append((fr, (os.path.basename(m.group(4)), int(m.group(5))) if m.group(3) else None))
        else:
            chain = [(fr, (os.path.basename(m.group(4)), int(m.group(5))) if m.group(3) else None)]
        chain_open = m.group(3) is not None
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        total += 1
        frames = [c[0] for c in chain] + ([chain[-1][1]] if chain and chain[-1][1] else [])
        name = "?"
        for f, ln in frames:
            fn = fn_of(f, ln)
            if f in ("hc_common.h", "hc_coop.h") or fn in HELPERS or not f.startswith("hc_"):
                continue
            name = f + ":" + fn
            break
        per_fn[name] += 1
print(f"{kernel}: {total} instructions = {total * 16 / 1024:.0f} KB")
shown = 0
for k, v in per_fn.most_common():
    if v * 16 < 4096:
        break
    shown += v
    print(f"{100 * v / total:5.1f}%  {v * 16 / 1024:7.1f} KB  {k}")
print(f"{100 * (total - shown) / total:5.1f}%  {(total - shown) * 16 / 1024:7.1f} KB  (functions below 4 KB)")
