"""Join an `ncu --page source --csv` SASS listing with the line table of the cubin that ran (nvdisasm -g) and
print, per source routine (file + enclosing function) and per source line, the share of executed
instructions and of the stall samples by reason.
usage: ncu_routines.py <source.csv> <cubin> <kernel substring> [top lines]"""
import csv, re, subprocess, sys, collections, os
rep_csv, cubin, func = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
inside = False; cur = ("?", 0); table = {}
for ln in out.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = func in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: table[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]
ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples")
reasons = ["stall_wait", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_no_inst", "stall_selected",
           "stall_branch_resolving", "stall_mio", "stall_math", "stall_dispatch"]
ri = {r: hdr.index(r) for r in reasons}
data = [r for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
base = int(data[0][0], 16)

def fn_ranges(path):
    res = []
    try:
        src = open(path).read().splitlines()
    except OSError:
        return res
    for i, l in enumerate(src, 1):
        if l.startswith((" ", "\t", "#", "//", "}")) or "(" not in l: continue
        m = re.search(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', l)
        if m and m.group(1) not in ("if", "for", "while", "switch", "defined", "__launch_bounds__", "static_assert"):
            res.append((i, m.group(1)))
        elif m and m.group(1) == "__launch_bounds__":
            res.append((i, "kernel"))
    return res
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cvxpnpl_b200", "csrc")
ranges = {f: fn_ranges(os.path.join(root, f)) for f in os.listdir(root)}
def bucket(f, l):
    r = ranges.get(f)
    if not r: return f
    name = "?"
    for s, n in r:
        if s <= l: name = n
        else: break
    return f"{f.split('.')[0].replace('pnpl_', '')}:{name}"
ci = collections.Counter(); cs = collections.Counter(); cr = collections.defaultdict(collections.Counter)
li = collections.Counter(); ls = collections.Counter(); lr = collections.defaultdict(collections.Counter)
for r in data:
    t = table.get(int(r[0], 16) - base, ("?", 0))
    b = bucket(*t)
    i = int(r[ie] or 0); s = int(r[ns] or 0)
    ci[b] += i; cs[b] += s; li[t] += i; ls[t] += s
    for k, idx in ri.items():
        v = int(r[idx] or 0)
        cr[b][k] += v; lr[t][k] += v
ti = sum(ci.values()); ts = sum(cs.values())
print(f"instructions executed {ti}  samples {ts}  static {len(data)}")
tot = collections.Counter()
for b in cr: tot.update(cr[b])
print("all: " + "  ".join(f"{k[6:]} {100 * tot[k] / ts:.1f}" for k in reasons))
print(f"{'routine':38s} inst%  smp%  | " + " ".join(f"{k[6:11]:>5s}" for k in reasons[:6]))
for b, s in cs.most_common(30):
    print(f"{b:38s} {100 * ci[b] / ti:5.1f} {100 * s / ts:5.1f}  | " + " ".join(f"{100 * cr[b][k] / ts:5.1f}" for k in reasons[:6]))
print("-- top lines by samples")
for t, s in ls.most_common(topn):
    print(f"{t[0]}:{t[1]:<5d} inst {100 * li[t] / ti:5.1f}% smp {100 * s / ts:5.1f}%  | " + " ".join(f"{100 * lr[t][k] / ts:5.1f}" for k in reasons[:6]))
