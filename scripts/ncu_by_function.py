"""Per-function / per-line aggregation of an ncu source page:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_by_function.py src.csv [csrc dir]"""
import bisect, collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "homotopycontinuation.jl_b200", "csrc")
cur_file = None; cur_line = None
per_line = collections.defaultdict(lambda: [0, 0, 0, 0.0])
ops = collections.Counter(); tot_inst = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; i_s = hdr.index('# Samples'); i_i = hdr.index('Instructions Executed'); i_l = hdr.index('stall_long_sb'); i_t = hdr.index('Thread Instructions Executed'); continue
    if r[0] != "": cur_line = (cur_file, int(r[0]), r[1].strip()[:100])
    if len(r) > i_l and r[2] not in ("", "..."):
        try: s = int(r[i_s]); n = int(r[i_i]); l = int(r[i_l]); t = int(r[i_t])
        except ValueError: continue
        v = per_line[cur_line]; v[0] += s; v[1] += n; v[2] += l; v[3] += t
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
        if m: ops[m.group(2)] += n
        tot_inst += n
tot = sum(v[0] for v in per_line.values()); toti = sum(v[1] for v in per_line.values()); tott = sum(v[3] for v in per_line.values())
print(f"samples {tot}  warp-inst {toti}  avg active threads {tott / max(toti, 1):.1f}")
print("opcode mix:", ", ".join(f"{k} {100 * n / tot_inst:.1f}%" for k, n in ops.most_common(18)))
def marks(path):
    out = []
    for i, l in enumerate(open(path).read().split('\n'), 1):
        m = re.match(r'\s*(template <[^>]*>\s*)?HC_(HDN|HD|D)\s+(static\s+)?[\w:<>&\*,\s]+?\s+(\w+)\(', l)
        if m: out.append((i, m.group(4)))
    return out
agg = collections.Counter(); aggi = collections.Counter()
for f in sorted({k[0] for k in per_line}):
    path = os.path.join(csrc, f)
    mk = marks(path) if os.path.exists(path) else []
    starts = [m[0] for m in mk]
    for (ff, ln, _), v in per_line.items():
        if ff != f: continue
        j = bisect.bisect_right(starts, ln) - 1
        name = f + ':' + (mk[j][1] if j >= 0 else '?')
        agg[name] += v[0]; aggi[name] += v[1]
print("--- by function")
for k, v in agg.most_common(40):
    print(f"{100 * v / tot:5.1f}% smp {100 * aggi[k] / toti:5.1f}% inst  {k}")
print("--- by line")
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{100 * v[0] / tot:5.2f}% smp {100 * v[1] / toti:5.2f}% inst lsb {100 * v[2] / max(v[0], 1):3.0f}%  {k[0]}:{k[1]}  {k[2]}")
