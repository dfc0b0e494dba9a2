"""Join an `ncu --page source --csv` SASS listing with the line table of the cubin that
ran (nvdisasm -g) and print the stall-sample share per source line / per function."""
import csv, re, subprocess, sys, collections
rep_csv, cubin, func = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# find function section
lines = out.splitlines()
inside = False; cur = ("?", 0); table = []   # (offset, file, line)
for ln in lines:
    if ln.startswith("//--------------------- .text."):
        inside = func in ln
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        table.append((int(m.group(1), 16), cur[0], cur[1], m.group(2)))
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]; ns = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
data = [(int(r[0], 16), int(r[ns] or 0), int(r[ie] or 0), r[1].strip()) for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
base = data[0][0]
off2 = {t[0]: t for t in table}
per_line = collections.Counter(); per_line_i = collections.Counter()
tot = sum(d[1] for d in data); toti = sum(d[2] for d in data)
miss = 0
for addr, smp, inst, sass in data:
    t = off2.get(addr - base)
    if t is None: miss += 1; key = ("?", 0)
    else: key = (t[1], t[2])
    per_line[key] += smp; per_line_i[key] += inst
print(f"samples {tot} instr {toti} unmatched {miss}/{len(data)}")
for key, smp in per_line.most_common(topn):
    print(f"{100*smp/tot:5.1f}% smp {100*per_line_i[key]/toti:5.1f}% inst  {key[0]}:{key[1]}")
