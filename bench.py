#!/usr/bin/env python
"""bench.py -- poses/sec on a batch of independent PnPL problems (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (assembly -> 10x10 SDP -> extraction) over
one synthetic batch of `--batch` problems per GPU (default 1e5 x PnPL with 8
points + 4 lines, fp64: BASELINE.json configs[2], the configuration the metric
is quoted on).  Weak scaling: every rank owns its own batch; the only collective
is the all-gather of the output poses.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poses/sec on 1e5-batch PnPL (8 pts + 4 lines)"
UNIT = "poses/s"
# algorithmic bytes per problem (SURVEY.md 8d / BASELINE.md): fp64 in 640 B, one pose out 96 B
BYTES_PER_PROBLEM = {(8, 4): 736, (8, 0): 416, (0, 6): 576}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=100_000, help="problems per GPU per step")
    ap.add_argument("--n-pts", type=int, default=8)
    ap.add_argument("--n-lines", type=int, default=4)
    ap.add_argument("--noise", type=float, default=1.0, help="pixel noise sigma (synth.py grid: 0,1,2)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--admm", default="f64", choices=["f64", "f32"],
                    help="f64: FP64 ADMM throughout (the headline, BASELINE.json configs[2]); f32: FP32 first "
                         "phase + FP64 tail and extraction (configs[3]), a side measurement")
    ap.add_argument("--large-n", type=int, default=0,
                    help="side measurement (not the headline): HBM roofline of the streaming assembly with this "
                         "many points per problem (benchmarks/scalability/pnp.py regime)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference path; oracle/) on the host cores
# ------------------------------------------------------------------------------------
def _cpu_worker(args):
    import warnings
    lo, hi, n_pts, n_lines, noise, seed = args
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    d = synth.make_batch(hi, n_pts, n_lines, noise=noise, seed=seed)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(lo, hi):
            if n_pts and n_lines:
                orc.pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"])
            elif n_pts:
                orc.pnp(d["pts_2d"][i], d["pts_3d"][i], d["K"])
            else:
                orc.pnl(d["line_2d"][i], d["line_3d"][i], d["K"])
    return time.perf_counter() - t0


def cpu_rate(sample, n_pts, n_lines, noise, cores, pool):
    """poses/s of the oracle path on `cores` processes over `sample` problems
    (reference defaults eps=1e-9, max_iters=2500; one call per problem, as the
    reference has no batch API)."""
    per = (sample + cores - 1) // cores
    jobs = [(k * per, min((k + 1) * per, sample), n_pts, n_lines, noise, 4242) for k in range(cores)]
    jobs = [j for j in jobs if j[1] > j[0]]
    t0 = time.perf_counter()
    pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return sample / wall, wall


def make_pool(cores):
    import multiprocessing as mp
    from oracle import scs_port
    scs_port.build()
    ctx = mp.get_context("fork")
    return ctx.Pool(cores)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_name(B, n_pts, n_lines, admm, noise):
    """config.workload, shared by both arms (the reference arm runs a bounded sample of it)."""
    kind = "PnPL" if (n_pts and n_lines) else ("PnP" if n_pts else "PnL")
    cfg_name = {(8, 4): "BASELINE.json configs[2]", (8, 0): "BASELINE.json configs[1]",
                (0, 6): "BASELINE.json configs[3]"}.get((n_pts, n_lines), "not a BASELINE.json config")
    return (f"{B} x {kind} ({n_pts} pts + {n_lines} lines) per GPU, "
            f"{'fp64' if admm == 'f64' else 'fp32 ADMM first phase + fp64'}, sigma={noise}px, Kinect K ({cfg_name})")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    pool = make_pool(cores)
    # bounded sample per step: about 3 s of wall clock on all cores (rate probed twice:
    # the first probe also pays for the pool's start-up)
    cpu_rate(max(2 * cores, 16), a.n_pts, a.n_lines, a.noise, cores, pool)
    r0, _ = cpu_rate(max(8 * cores, 64), a.n_pts, a.n_lines, a.noise, cores, pool)
    sample = a.cpu_sample or max(8 * cores, int(r0 * 3.0))
    for _ in range(a.warmup):
        cpu_rate(max(cores, sample // 4), a.n_pts, a.n_lines, a.noise, cores, pool)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_rate(sample, a.n_pts, a.n_lines, a.noise, cores, pool)
    wall = time.perf_counter() - t0
    pool.close()
    value = a.steps * sample / wall
    desc = (f"{sample} problems per step of the same synthetic PnPL workload, one oracle call per problem "
            f"(numpy restatement of cvxpnpl.py + oracle/scs_port.c, eps=1e-9, max_iters=2500), "
            f"{cores} processes")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a.batch, a.n_pts, a.n_lines, "f64", a.noise),
                   "problems_per_gpu_per_step": a.batch, "eps": 1e-9, "max_iters": 2500,
                   "sample_problems_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region.  The sampler is started -- and its
    first sample awaited -- before the warm-up (nvidia-smi's start-up takes driver locks for a few hundred
    ms and, started right at the timed region, it delayed one of ten steps by 4 ms in one run); only the
    samples whose timestamp falls inside the timed region are kept."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(self.index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
            return
        t = time.time()
        while time.time() - t < 5.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[2]), float(c[3]),
                             [nm for nm, v in zip(names, c[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        where = "timed region"
        if not inside:   # clock skew between nvidia-smi's timestamps and time.time(): keep everything
            inside, where = rows, "warm-up + timed region"
        sm = [r[1] for r in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": inside[-1][2] if inside else None,
                "reasons": sorted({x for r in inside for x in r[3]}), "samples": len(sm), "window": where}


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import cvxpnpl_b200 as cb
    from cvxpnpl_b200 import synth
    from cvxpnpl_b200.distributed import RECORD, gather_records, pack_record

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, n_pts, n_lines = a.batch, a.n_pts, a.n_lines

    # synthetic workload (restated benchmarks/toolkit/suites/synth.py), one shard per rank
    d = synth.make_batch(B, n_pts, n_lines, noise=a.noise, seed=42 + rank)
    host = {k: torch.from_numpy(d[k]).pin_memory() for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
    K = torch.from_numpy(d["K"]).to(dev)
    devin = {k: v.to(dev) for k, v in host.items()}
    ws = cb.Workspace(B, dev)
    out = None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    h2d_bytes = sum(v.numel() * 8 for v in host.values())
    host_out = torch.empty((B, RECORD), dtype=torch.float64).pin_memory()
    d2h_bytes = host_out.numel() * 8

    def kernel_step(inp):
        nonlocal out
        out = cb.solve_batched(K, pts_2d=inp["pts_2d"] if n_pts else None, pts_3d=inp["pts_3d"] if n_pts else None,
                               line_2d=inp["line_2d"] if n_lines else None,
                               line_3d=inp["line_3d"] if n_lines else None, workspace=ws, out=out,
                               admm_dtype=a.admm)
        return out

    def pack(o):
        return pack_record(o.R[:, 0], o.t[:, 0], o.n_poses, o.status, o.iters)

    def step_device():
        o = kernel_step(devin)
        if world > 1:
            gather_records(pack(o), world * B)   # the one collective: all-gather of the poses
        return o

    # end to end through the public host-memory entry point: cvxpnpl_b200.HostStager copies the pinned
    # correspondences up in 4 slices (copy stream) and runs the pre-pass of each slice under the copies of
    # the next ones; then the solve, the record packing, the all-gather and the D2H read of the records
    stager = cb.HostStager(K, dev, chunks=4, admm_dtype=a.admm)
    hostd = {k: v for k, v in host.items() if (n_pts if k.startswith("pts") else n_lines)}

    def step_e2e():
        o = stager.solve(hostd)
        p = pack(o)
        if world > 1:
            gather_records(p, world * B)
        host_out.copy_(p, non_blocking=True)
        return o

    def timed(fn, steps, warmup, marks=None):
        for _ in range(warmup):
            fn()
            flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if marks is not None:
            marks.mark_begin()
        evs = []
        for _ in range(steps):
            flush.zero_()  # flush L2 between timed iterations (inputs 64 MB < 126 MB L2)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if marks is not None:
            marks.mark_end()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, a.steps, a.warmup, marks=sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None

    # kernel-only times for the roofline: CUDA events recorded by the library between its
    # launches, on the stream they are launched on (torch's current stream), over a
    # second pass of `steps` L2-flushed steps
    ksum, ms_kernel = {}, 0.0
    for _ in range(a.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = cb.solve_batched(K, pts_2d=devin["pts_2d"] if n_pts else None, pts_3d=devin["pts_3d"] if n_pts else None,
                               line_2d=devin["line_2d"] if n_lines else None,
                               line_3d=devin["line_3d"] if n_lines else None, workspace=ws, out=out,
                               admm_dtype=a.admm, timing=True)
        e.record()
        for k, v in cb.last_kernel_times().items():
            ksum[k] = ksum.get(k, 0.0) + v
        ms_kernel += s.elapsed_time(e)
    ms_kernel /= a.steps
    kernel_ms = {k: v / a.steps for k, v in ksum.items() if v > 0}
    dominant = max(kernel_ms, key=kernel_ms.get)
    launches_per_step = out.launches

    # side measurement (not the headline): "fp32 ADMM + fp64 extraction" (BASELINE.json
    # configs[3] precision mode) on the same batch, device-resident inputs
    ms_mixed = None
    if a.admm == "f64":
        def step_mixed():
            nonlocal out
            out = cb.solve_batched(K, pts_2d=devin["pts_2d"] if n_pts else None,
                                   pts_3d=devin["pts_3d"] if n_pts else None,
                                   line_2d=devin["line_2d"] if n_lines else None,
                                   line_3d=devin["line_3d"] if n_lines else None, workspace=ws, out=out,
                                   admm_dtype="f32")
        ms_mixed = timed(step_mixed, a.steps, a.warmup)
        kernel_step(devin)   # leave the FP64 result in `out` for the quality block

    ms_e2e = timed(step_e2e, a.steps, a.warmup)

    # streaming variant (extra, not the e2e headline): cvxpnpl_b200.HostPipeline double-buffers the
    # inputs, so the H2D copy of step k+1 runs on a copy stream during the solve of step k.  Every
    # one of the K input batches is still copied inside the timed region (the first one up front).
    ms_pipe = None
    if world == 1:
        pipe = cb.HostPipeline(K, dev, admm_dtype=a.admm)
        calls = {"n": 0, "total": a.warmup + a.steps}

        def step_pipe():
            calls["n"] += 1
            pipe.step(hostd, hostd if calls["n"] < calls["total"] else None)
        ms_pipe = timed(step_pipe, a.steps, a.warmup)

    # health of the result (not timed): status histogram, iterations, error vs ground truth
    torch.cuda.synchronize()
    st = (out.status & 0xFF).cpu().numpy()
    iters = out.iters.cpu().numpy()
    ang, terr = synth.pose_error(d["R_gt"], d["t_gt"], out.R[:, 0].cpu().numpy(), out.t[:, 0].cpu().numpy())

    fp64_peak = cb.measure_fp64_peak(dev) if rank == 0 else None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        kc = {}
        kc_path = os.path.join(ROOT, "profiles", "r1_kernel_constants.json")
        if os.path.exists(kc_path) and (n_pts, n_lines, B) == (8, 4, 100_000):
            with open(kc_path) as f:
                kc = json.load(f)
        total = world * B * a.steps
        value = total / (ms_dev * 1e-3)
        e2e = total / (ms_e2e * 1e-3)
        bpp = BYTES_PER_PROBLEM.get((n_pts, n_lines), 8 * (5 * n_pts + 10 * n_lines) + 96)
        achieved = bpp * B / (kernel_ms[dominant] * 1e-3) / 1e9
        flops = kc.get("fp64_flops_per_launch")
        if flops and kc.get("iters_mean_at_capture"):
            # the count is per launch of the captured build; FP64 work is proportional to the DR
            # iterations, so a batch that needs fewer / more of them is scaled accordingly
            flops = flops * float(iters.mean()) / kc["iters_mean_at_capture"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if a.admm == "f64" else "f32 first phase + f64 tail/extraction", "data": "synthetic",
            "config": {"workload": workload_name(B, n_pts, n_lines, a.admm, a.noise),
                       "problems_per_gpu_per_step": B, "eps": 1e-9, "max_iters": 2500,
                       "l2": "flushed between timed iterations (256 MB write)",
                       "collective": "all_gather of [B,15] pose records (NCCL)" if world > 1 else "none"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / a.steps},
            "e2e_pipelined": ({"what": "same as e2e but through cvxpnpl_b200.HostPipeline: the H2D copy of the next "
                                       "batch overlaps the solve of the current one (copy stream); extra, not the headline",
                               "value": total / (ms_pipe * 1e-3), "unit": UNIT,
                               "ms_per_step": ms_pipe / a.steps} if ms_pipe else None),
            "gpu_launches": launches_per_step * a.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": (kc["dram_bytes_read"] + kc["dram_bytes_write"]) if kc else None,
                         "peak_kind": peak_kind,
                         "kernel": dominant, "kernel_ms": kernel_ms[dominant],
                         "all_kernels_ms": kernel_ms, "step_ms": ms_kernel,
                         "algorithmic_bytes_per_problem": bpp,
                         "note": "the path is compute/latency bound in shared memory + fp64 pipe, not HBM bound "
                                 "(SURVEY.md 8d): the HBM fraction is reported as asked, see DESIGN.md"},
            # what actually bounds the kernel: the FP64 pipe + shared memory, fed by one warp
            # per scheduler.  FP64 flops of the dominant kernel per launch counted by ncu
            # (profiles/r1_kernel_constants.json), over its measured duration here
            "fp64": {"peak_tflops_measured": fp64_peak, "flops_per_launch": flops,
                     "achieved_tflops": (flops / (kernel_ms[dominant] * 1e-3) / 1e12) if flops else None,
                     "frac": (flops / (kernel_ms[dominant] * 1e-3) / 1e12 / fp64_peak) if flops else None},
            "mixed_precision_side": ({"what": "fp32 ADMM first phase + fp64 tail and extraction (configs[3] mode) on "
                                              "the same batch; not the headline",
                                      "value": total / (ms_mixed * 1e-3), "unit": UNIT,
                                      "ms_per_step": ms_mixed / a.steps} if ms_mixed else None),
            "quality": {"status_hist": np.bincount(st, minlength=5).tolist(),
                        "iters_median": float(np.median(iters)), "iters_p99": float(np.percentile(iters, 99)),
                        "iters_max": int(iters.max()), "rot_err_vs_gt_median_rad": float(np.nanmedian(ang)),
                        "t_err_vs_gt_median": float(np.nanmedian(terr))},
        }
        if world == 1 and not a.no_cpu_baseline:
            cores = host_cores()
            pool = make_pool(cores)
            cpu_rate(max(2 * cores, 16), n_pts, n_lines, a.noise, cores, pool)
            r0, _ = cpu_rate(max(8 * cores, 64), n_pts, n_lines, a.noise, cores, pool)
            sample = a.cpu_sample or max(8 * cores, int(r0 * 15.0))  # about 15 s of CPU work
            v, wall = cpu_rate(sample, n_pts, n_lines, a.noise, cores, pool)
            pool.close()
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{sample} problems of the same workload in {wall:.1f} s, one oracle call per problem "
                          f"(numpy restatement of cvxpnpl.py + oracle/scs_port.c standing in for SCS, eps=1e-9, "
                          f"max_iters=2500), {cores} processes"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_large_n(a):
    """Achieved HBM bandwidth of the large-n assembly (40 B per point streamed once)."""
    import torch
    import cvxpnpl_b200 as cb
    from cvxpnpl_b200 import synth
    n, B = a.large_n, max(1, min(65535, int(2e9 // (40 * a.large_n))))   # ~2 GB of correspondences (> L2)
    dev = torch.device("cuda", 0)
    d = synth.make_batch(B, n, 0, noise=1.0, seed=1)
    p2, p3, K = (torch.from_numpy(d[k]).to(dev) for k in ("pts_2d", "pts_3d", "K"))
    for _ in range(3):
        cb.assemble_batched(K, p2, p3)
    torch.cuda.synchronize()
    evs = []
    for _ in range(a.steps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        cb.assemble_batched(K, p2, p3)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / a.steps
    peaks, kind = measured_peaks()
    gbs = 40.0 * n * B / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "large-n assembly", "points_per_problem": n, "problems": B, "ms": ms,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": gbs / peaks["hbm_gbs"], "peak_kind": kind,
                                   "algorithmic_bytes_per_point": 40}}))


if __name__ == "__main__":
    args = parse()
    if args.large_n:
        run_large_n(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
