"""Dev tool: profiles/r<N>_kernel_constants.json from an ncu report of tools/prof_run.py.
    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    ncu -i rep.ncu-rep --page source --csv > src.csv          (report captured with -k regex:<kernel>)
    python tools/kernel_constants.py raw.csv src.csv "<source note>" <mean iterations> [ncu kernel name [kernel_times slot name]] > profiles/r2_kernel_constants.json
Thread-level FP64 flops of the solver kernel = 2 DFMA + DMUL + DADD, counted per SASS line
("Predicated-On Thread Instructions Executed")."""
import csv
import json
import sys

raw, src, note = sys.argv[1], sys.argv[2], sys.argv[3]
iters_mean = float(sys.argv[4]) if len(sys.argv) > 4 else None   # mean DR iterations of the captured batch
kname = sys.argv[5] if len(sys.argv) > 5 else None
label = sys.argv[6] if len(sys.argv) > 6 else kname   # name of the kernel's slot in cvxpnpl_b200.last_kernel_times()
rows = list(csv.reader(open(raw)))
hdr = rows[0]
r = next(x for x in rows[2:] if (kname in x[hdr.index("Kernel Name")] if kname else
                                ("solve_fused_kernel<0>" in x[hdr.index("Kernel Name")] or "solve_fused_kernel<(bool)0>" in x[hdr.index("Kernel Name")])))
g = lambda n: float(r[hdr.index(n)])
unit = rows[1][hdr.index("dram__bytes_read.sum")]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
cnt = {"DFMA": 0, "DMUL": 0, "DADD": 0}
first = True
for x in csv.reader(open(src)):
    if x and x[0] == "Kernel Name":
        if not first:   # the page repeats every kernel; the first table is the one to count
            break
        first = False
        continue
    if x and x[0] == "Address":
        h = x
        continue
    if len(x) > 8 and x[0].isdigit() or (x and x[0].startswith("0x")):
        op = x[1].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        k = op.split(".")[0]
        if k in cnt:
            cnt[k] += int(x[h.index("Predicated-On Thread Instructions Executed")] or 0)
flops = 2 * cnt["DFMA"] + cnt["DMUL"] + cnt["DADD"]
print(json.dumps({
    "source": note,
    "kernel": label or "solve_fused_kernel",
    "ncu_kernel_name": kname,
    "dram_bytes_read": int(g("dram__bytes_read.sum") * scale),
    "dram_bytes_write": int(g("dram__bytes_write.sum") * scale),
    "warp_inst_executed": int(g("smsp__inst_executed.sum")),
    "thread_inst_dfma": cnt["DFMA"], "thread_inst_dmul": cnt["DMUL"], "thread_inst_dadd": cnt["DADD"],
    "fp64_flops_per_launch": flops,
    "kernel_ms_under_ncu": round(g("gpu__time_duration.sum"), 3),
    "iters_mean_at_capture": iters_mean,
    "note": "FP64 flops (2 DFMA + DMUL + DADD, thread level) of the persistent FP64 solver kernel only; the other kernels of "
            "the step are not counted.  DRAM traffic of this kernel: the pre-pass record of every problem is read, the parked "
            "state written.  iters_mean_at_capture: mean DR iterations of the captured batch (tools/prof_run.py); bench.py "
            "scales the flop count by the mean iterations it measures.",
}, indent=1))
