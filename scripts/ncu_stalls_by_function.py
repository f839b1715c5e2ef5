"""Per-function stall-reason breakdown of an ncu source page (SASS view with inline chains is not needed):
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_stalls_by_function.py src.csv
Charges the samples of a CUDA-C line to the source function that contains it (same attribution as ncu_by_function.py)."""
import bisect, collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "homotopycontinuation.jl_b200", "csrc")
marks = {}
def fn_of(f, ln):
    if f not in marks:
        out = []
        try:
            for i, l in enumerate(open(os.path.join(csrc, f)).read().split('\n'), 1):
                m = re.match(r'\s*(template <[^>]*>\s*)?(static\s+)?HC_(HDN|HD|D)\s+(static\s+)?[\w:<>&\*,\s]+?\s+(\w+)\(', l)
                if m: out.append((i, m.group(5)))
        except OSError:
            pass
        marks[f] = out
    mk = marks[f]
    k = bisect.bisect_right([a for a, _ in mk], ln) - 1
    return mk[k][1] if k >= 0 else "?"
cur_file = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r
        cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        i_s = hdr.index('# Samples'); i_i = hdr.index('Instructions Executed')
        continue
    if r[0] == "" or cur_file is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    key = cur_file + ":" + fn_of(cur_file, ln)
    try:
        s = int(r[i_s] or 0)
    except ValueError:
        continue
    agg[key]["samples"] += s; tot["samples"] += s
    try: agg[key]["inst"] += int(r[i_i] or 0); tot["inst"] += int(r[i_i] or 0)
    except ValueError: pass
    for h, i in cols.items():
        try: v = int(r[i] or 0)
        except ValueError: v = 0
        agg[key][h] += v; tot[h] += v
names = ["stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_math", "stall_lg", "stall_barrier"]
print("total samples", tot["samples"], " ".join(f"{n[6:]} {100 * tot[n] / max(1, tot['samples']):.1f}%" for n in names))
print(f"{'function':44s} {'smp%':>6s} {'inst%':>6s} " + " ".join(f"{n[6:14]:>8s}" for n in names))
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    print(f"{key:44s} {100 * c['samples'] / tot['samples']:6.1f} {100 * c['inst'] / max(1, tot['inst']):6.1f} " +
          " ".join(f"{100 * c[n] / max(1, c['samples']):8.0f}" for n in names))
