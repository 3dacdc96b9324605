#!/usr/bin/env python
"""Dev tool (GPU box): the opt-in exact-pruned ADD-S kernel (b') against the all-pairs kernel (b):
bit equality of every output byte and both rates, on config-2-style workloads at several mesh sizes
and perturbation levels.   python tools/pruned_ab.py [poses_per_case]"""
import importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
core, W = pkg.core, pkg.workloads
dev = torch.device("cuda", 0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b) * 1e-3)
    return best

g = torch.Generator(device=dev); g.manual_seed(11)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
out = []
for npts, mesh in ((2048, "sphere"), (2048, "box"), (1000, "sphere"), (500, "sphere"), (500, "box"), (131, "sphere"), (37, "sphere")):
    for sigma in (0.02, 0.1, 0.5):
        mk = (lambda n, s: W.sphere_mesh(n, 0.1646, s)) if mesh == "sphere" else (lambda n, s: W.box_mesh(n, (0.1, 0.12, 0.05), s))
        pts = {9: mk(npts, 200 + npts), 10: mk(npts, 300 + npts)}
        table = core.MeshTable(pts, {9: 0.1646, 10: 0.1759}, pkg.SYMMETRIC_OBJECT_IDS, dev)
        B = m if npts >= 1000 else 4 * m
        obj = torch.where(torch.rand(B, generator=g, device=dev) < 0.5, 9, 10).to(torch.int64)
        qa, ta = torch.nn.functional.normalize(rnd(B, 4), dim=1), rnd(B, 3) * 0.1 + torch.tensor([0.0, 0.0, 0.8], device=dev)
        qb, tb = torch.nn.functional.normalize(qa + sigma * rnd(B, 4), dim=1), ta + 0.05 * sigma * rnd(B, 3)
        qb[5] = float("nan"); tb[6, 1] = float("inf"); qb[7] = qa[7]; tb[7] = ta[7]; qb[8] *= 3.0; obj[9] = 4
        order = torch.argsort(obj, stable=True).to(torch.int32)
        full = table.evaluate(qb, tb, qa, ta, obj, order=order)[4]
        prun = table.evaluate(qb, tb, qa, ta, obj, order=order, prune=True)[4]
        same = bool(torch.equal(full[:10 * B], prun[:10 * B]))
        t_full = timed(lambda: table.evaluate_packed(qb, tb, qa, ta, obj, order=order))
        t_pr = timed(lambda: table.evaluate_packed(qb, tb, qa, ta, obj, order=order, prune=True))
        r = {"n": npts, "mesh": mesh, "sigma": sigma, "poses": B, "bits_equal": same, "all_pairs_Mposes_s": B / t_full / 1e6,
             "pruned_Mposes_s": B / t_pr / 1e6, "speedup": t_full / t_pr}
        if not same:
            a = full[:4 * B].view(torch.float32); b = prun[:4 * B].view(torch.float32)
            a2 = full[4 * B:8 * B].view(torch.float32); b2 = prun[4 * B:8 * B].view(torch.float32)
            bad = torch.nonzero(a2.view(torch.int32) != b2.view(torch.int32)).flatten()
            r["add_diff"] = int((a.view(torch.int32) != b.view(torch.int32)).sum()); r["adds_diff"] = int(bad.numel())
            r["first_bad"] = [(int(i), float(a2[i]), float(b2[i])) for i in bad[:5]]
        print(json.dumps(r), flush=True)
        out.append(r)
print(json.dumps({"all_equal": all(r["bits_equal"] for r in out)}))
