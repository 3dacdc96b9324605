#!/usr/bin/env python
"""bench.py -- ADD-S pose evaluations per second on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic pose hypotheses:
BASELINE config 2, i.e. 65,536 hypotheses against 2,048-point symmetric-object meshes
(ADD + ADD-S + ADD-0.1d decision per pose, one kernel launch) per GPU, followed at N > 1
by the NCCL all-reduce of the per-object hit counts.  Weak scaling: every rank evaluates
its own 65,536 hypotheses; `value` is the whole-job rate.

  python bench.py [--gpus N --steps K --warmup W]        our arm (one process per GPU)
  python bench.py --impl reference [...]                 CPU arm: the oracle port of the
        reference's algorithm on all host cores, same metric, bounded sample per step

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the roofline arithmetic.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_POINTS = 2048
POSES_PER_GPU = 65536
FLOP_PER_POSE = 8 * N_POINTS * N_POINTS            # 3 sub + 3 mul + 2 add per pair
BYTES_PER_POSE = 64 + 10                            # 14 floats + int64 id in, 2 floats + 2 bytes out
METRIC = "ADD-S pose evals/sec (2k-pt mesh)"
UNIT = "poses/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--poses", type=int, default=POSES_PER_GPU, help="hypotheses per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="poses timed on the CPU baseline (0 = sized for ~12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the HBM-bound secondary kernels")
    return ap.parse_args()


def config_dict(args, extra=None):
    c = {"workload": f"BASELINE config 2: ADD-S symmetric eval, {N_POINTS}-point synthetic meshes (ids 9,10), "
                     f"{args.poses} pose hypotheses per GPU per step",
         "n_points": N_POINTS, "poses_per_gpu_per_step": args.poses,
         "l2": "inputs (4.8 MB) are smaller than L2; a 256 MiB buffer is written between timed steps to flush it "
               "(the kernel is FP32-issue-bound, DRAM traffic is ~0.1% of the time)",
         "parallelism": f"hypothesis-sharded x{args.gpus}, NCCL all-reduce of int64 hit counts only"}
    if extra:
        c.update(extra)
    return c


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the ADD-S kernel, from the newest
    committed `ncu --set full` summary under profiles/ (same workload and launch shape as the bench)."""
    import glob, re
    best = None
    for path in sorted(glob.glob(os.path.join(REPO, "profiles", "adds_*_summary.txt"))):
        txt = open(path).read()
        r = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", txt)
        w = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", txt)
        if not r or not w:
            continue
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        best = {"bytes": float(r.group(1)) * unit.get(r.group(2), 1.0) + float(w.group(1)) * unit.get(w.group(2), 1.0),
                "source": os.path.relpath(path, REPO)}
    return best


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [l for (t, l) in self.lines if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.05)]
        if not inside:
            inside = [l for (_, l) in self.lines[-3:]]
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- CPU arm
def cpu_rate(W, n_poses, threads=None):
    """Oracle port of the reference algorithm on the host cores (bounded sample)."""
    import oracle as O
    O.build()
    threads = threads or O.max_threads()
    pts, dia = W.config2_meshes(N_POINTS)
    pq, pt, gq, gt, obj = W.config2(n_poses)
    table = O.MeshTable(pts, dia)
    O.add_eval(table, pq[:threads], pt[:threads], gq[:threads], gt[:threads], obj[:threads], n_threads=threads)
    t0 = time.perf_counter()
    O.add_eval(table, pq, pt, gq, gt, obj, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_poses / dt, threads, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, kind 'port': the
    reference is Python/PyTorch and cannot be installed as a package -- no setup.py; the
    port was validated bit-for-bit against it, tests/golden) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg_w = importlib.import_module("6d-pose-estimation_b200.workloads")
    import oracle as O
    O.build()
    threads = O.max_threads()
    pts, dia = pkg_w.config2_meshes(N_POINTS)
    table = O.MeshTable(pts, dia)
    # one step = a bounded sample of the workload: ~2-4 s of CPU work
    probe, _, _ = cpu_rate(pkg_w, max(64, 8 * threads), threads)
    sample = int(max(64, min(args.poses, probe * 3.0)))
    pq, pt, gq, gt, obj = pkg_w.config2(sample)
    for _ in range(args.warmup):
        O.add_eval(table, pq[:sample // 4], pt[:sample // 4], gq[:sample // 4], gt[:sample // 4],
                   obj[:sample // 4], n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.add_eval(table, pq, pt, gq, gt, obj, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, {"sample_poses_per_step": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} poses of config 2 per step x {args.steps} steps; oracle/pose_oracle.c on all "
                                       "host threads (bit-exact restatement; the reference's own PyTorch-eager loop measured "
                                       "46.5 poses/s on 8 cores for this mesh size, BASELINE.md section 2)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def secondary_rooflines(pkg, dev, hbm_peak):
    """Kernels (a), (c), (d1), (d2) at scaled batches, timed with CUDA events around the bare C-ABI
    call (no Python wrapper in the timed region).  Algorithmic bytes per row: DESIGN.md section 4."""
    import torch
    core, W = pkg.core, pkg.workloads
    L, st = core.lib(), core.stream_ptr(dev)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        return best

    out = []
    n = 1 << 22
    g = torch.Generator(device=dev); g.manual_seed(7)
    rnd = lambda *shape: torch.randn(*shape, generator=g, device=dev)
    pq, gq, pt, gt = rnd(n, 4), rnd(n, 4), rnd(n, 3), rnd(n, 3)
    o3 = torch.empty(3, device=dev); g1 = torch.empty_like(pq); g2 = torch.empty_like(pt)
    ws = torch.zeros(64, dtype=torch.uint8, device=dev)
    t = timed(lambda: core.check(L.p6d_pose_loss_fwd_bwd(pq.data_ptr(), pt.data_ptr(), gq.data_ptr(), gt.data_ptr(), n, 1.0,
                                                         10.0, 0, o3.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                                         ws.data_ptr(), dev.index, st)))
    out.append(("pose_loss_fwd_bwd (c)", n, 84, t))
    z, uv = torch.rand(n, device=dev) + 0.4, torch.rand(n, 2, device=dev) * 400
    K = torch.tensor(pkg.DEFAULT_K, dtype=torch.float32, device=dev).expand(n, 3, 3).contiguous()
    o = torch.empty(n, 3, device=dev)
    t = timed(lambda: core.check(L.p6d_pinhole_fwd(z.data_ptr(), uv.data_ptr(), K.data_ptr(), 1, n, o.data_ptr(), dev.index, st)))
    out.append(("pinhole_fwd (d1)", n, 60, t))
    m = 1 << 20
    d8 = torch.rand(m, 8, 8, device=dev) * 1.5
    uv8 = torch.rand(m, 2, device=dev) * 8
    o8 = torch.empty(m, 3, device=dev)
    t = timed(lambda: core.check(L.p6d_depth_backproject(d8.data_ptr(), 8, 8, uv8.data_ptr(), K.data_ptr(), 1, m, 7.0,
                                                         o8.data_ptr(), dev.index, st)))
    out.append(("depth_backproject (d2), 8x8 crops (one 32-B sector of each 256-B crop is read)", m, 32 + 8 + 36 + 12, t))
    res = [{"kernel": k, "bound": "hbm", "rows": rows, "bytes_per_row": bpr, "us": round(t * 1e6, 1),
            "achieved": rows * bpr / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": rows * bpr / t / 1e9 / hbm_peak if hbm_peak else None} for k, rows, bpr, t in out]
    # (a) ADD only is FP32-issue-bound, not HBM-bound: report poses/s
    pts = {0: W.sphere_mesh(1000, 0.102, 100)}
    table = core.MeshTable(pts, {0: 0.102}, pkg.SYMMETRIC_OBJECT_IDS, dev)
    obj = torch.zeros(m, dtype=torch.int64, device=dev)
    qa, ta = torch.nn.functional.normalize(rnd(m, 4), dim=1), rnd(m, 3)
    t = timed(lambda: table.evaluate(qa, ta, qa, ta, obj, want_adds=False), 5)
    res.append({"kernel": "add_warp_kernel (a), N=1000", "bound": "fp32-issue", "poses": m, "us": round(t * 1e6, 1),
                "poses_per_s": m / t, "hbm_gbs": m * 73 / t / 1e9})
    return res


# ------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keeps stdout to the one JSON line of the contract
        dist.init_process_group("nccl", device_id=dev)

    pkg = importlib.import_module("6d-pose-estimation_b200")
    core, W = pkg.core, pkg.workloads
    info = core.device_info(local)

    B = args.poses
    pts, dia = W.config2_meshes(N_POINTS)
    table = core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, dev)
    # each rank gets its own seeded hypotheses (weak scaling)
    host = W.config2(B, base_seed=2000 + 1000 * rank)
    pinned = [torch.from_numpy(x).pin_memory() for x in host]
    d_in = [t.to(dev) for t in pinned]
    order = torch.argsort(d_in[4], stable=True).to(torch.int32)
    acc = [torch.zeros(table.n_slots, dtype=torch.int64, device=dev) for _ in range(2)] + \
          [torch.zeros(table.n_slots, dtype=torch.float64, device=dev) for _ in range(2)]
    counts = torch.zeros(2, table.n_slots, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches = 0

    def step():
        nonlocal launches
        for a in acc:
            a.zero_()
        table.evaluate(*d_in, want_adds=True, order=order, acc=acc)
        launches += 1
        if world > 1:
            counts[0].copy_(acc[0]); counts[1].copy_(acc[1])
            dist.all_reduce(counts)          # per-object ADD-0.1d hit / valid counts: the only exchange

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                # L2 flush between timed steps (outside the event pair)
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    total_ms = float(dev_ms.item())
    value = world * B * args.steps / (total_ms * 1e-3)
    timed_launches = launches

    # result sanity on the numbers just computed (accuracy is counts / totals, integer-exact)
    hits = int(acc[0].sum().item()); valid = int(acc[1].sum().item())
    assert valid == B, (valid, B)

    # ---------------- end-to-end: host buffers in, host results out, through the C ABI
    h_np = [t.numpy() for t in pinned]
    for _ in range(2):
        table.evaluate_host(*h_np, want_adds=True, per_pose=True)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        r = table.evaluate_host(*h_np, want_adds=True, per_pose=True)
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(e2e_dt.item())
    assert int(r["obj_hits"].sum()) == hits, "host-entry decisions differ from the device entry"

    if rank == 0:
        kernel_ms = float(np.mean(step_ms))            # rank 0's kernel; at N = 1 the step is the kernel
        achieved = B * FLOP_PER_POSE / (kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        nominal = info["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
        tr = profiled_traffic() if B == POSES_PER_GPU else None
        relaid = bool(core.lib().p6d_adds_schedule())
        roof = {"bound": "fp32", "achieved": achieved, "peak": nominal, "unit": "TFLOP/s",
                "kernel": ("adds_cta_kernel<256,8,2,deferred minima>, scan loop re-laid after linking by "
                           "csrc/sass_sched.py") if relaid else "adds_cta_kernel<512,4,2,2> (ptxas schedule)",
                "frac": achieved / nominal, "traffic": tr["bytes"] if tr else None,
                "traffic_source": tr["source"] if tr else None,
                "algorithmic_bytes_per_launch": B * BYTES_PER_POSE,
                "peak_source": f"nominal FFMA peak {info['sm_count']} SM x 128 lanes x 2 FLOP x {sm_max:.0f} MHz "
                               "(sm_max_mhz of MEASURED_PEAKS.json; that file has no FP32 entry -- the kernel "
                               "is neither HBM- nor tensor-bound, SURVEY 7.3.2)",
                "algorithmic_flop_per_pose": FLOP_PER_POSE,
                "hbm": {"achieved_gbs": B * BYTES_PER_POSE / (kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs"), "note": "of measured; informational"}}
        if not args.no_microbench:
            mb = {}
            for kind, name in ((0, "ffma"), (1, "ffma2"), (2, "adds_mix")):
                mb[name] = round(max(core.fp32_microbench(kind, local, 4000)[0] for _ in range(3)), 2)
            roof["measured_fp32_tflops"] = mb
            roof["frac_of_measured_mix"] = achieved / mb["adds_mix"] if mb["adds_mix"] else None
            roof["measured_mix_note"] = ("adds_mix = the scan tile on register operands as ptxas schedules it; the "
                                         "product loop is re-laid after linking, so a ratio above 1 is expected")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args), "roofline": roof, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r["h2d_bytes"],
                        "d2h_bytes_per_step": r["d2h_bytes"], "steps": e2e_steps,
                        "api": "p6d_add_eval_host via MeshTable.evaluate_host (pinned host buffers)"},
                "gpu_launches": timed_launches, "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
                "add_01d_acc": 100.0 * hits / valid}
        if not args.no_secondary and B == POSES_PER_GPU:
            try:
                line["secondary_rooflines"] = secondary_rooflines(pkg, dev, peaks.get("hbm_gbs"))
            except Exception as e:  # never lose the headline line over the side measurements
                line["secondary_rooflines"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            n_cpu = args.cpu_sample
            if n_cpu <= 0:
                probe, _, _ = cpu_rate(W, 256)
                n_cpu = int(min(B, max(256, probe * 12.0)))
            v, cores, dt = cpu_rate(W, n_cpu)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n_cpu} poses of the same config-2 workload, {dt:.1f} s, "
                                              "oracle/pose_oracle.c (C restatement, bit-exact vs the reference)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract goes to the real stdout; everything else any library
    prints during the run (e.g. NCCL's version banner) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
