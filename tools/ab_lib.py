#!/usr/bin/env python
"""Dev tool (GPU box): time kernels (a) and (c) of several builds of libp6d.so side by side.
   python tools/ab_lib.py LIB [LIB ...]      each LIB is loaded in a fresh process."""
import importlib, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

def one(lib):
    import numpy as np, torch
    pkg = importlib.import_module("6d-pose-estimation_b200")
    core, W = pkg.core, pkg.workloads
    core.SO_PATH = os.path.abspath(lib)
    import oracle as O
    dev = torch.device("cuda", 0)
    L, st = core.lib(), core.stream_ptr(dev)
    def timed(fn, reps=8):
        for _ in range(3): fn()
        torch.cuda.synchronize(); best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(200_000); a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b) * 1e-3)
        return best
    out = {"lib": lib}
    g = torch.Generator(device=dev); g.manual_seed(7)
    rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
    m = 1 << 20
    for n in (500, 1000, 2048):
        pts = {0: W.sphere_mesh(n, 0.102, 100)}
        table = core.MeshTable(pts, {0: 0.102}, pkg.SYMMETRIC_OBJECT_IDS, dev)
        obj = torch.zeros(m, dtype=torch.int64, device=dev)
        qa, ta = torch.nn.functional.normalize(rnd(m, 4), dim=1), rnd(m, 3)
        qb, tb = torch.nn.functional.normalize(qa + 0.05 * rnd(m, 4), dim=1), ta + 0.005 * rnd(m, 3)
        t = timed(lambda: table.evaluate(qb, tb, qa, ta, obj, want_adds=False))
        add = table.evaluate(qb, tb, qa, ta, obj, want_adds=False)[0][:2048].cpu().numpy()
        ref = O.add_eval(O.MeshTable(pts, {0: 0.102}), qb[:2048].cpu().numpy(), tb[:2048].cpu().numpy(), qa[:2048].cpu().numpy(),
                         ta[:2048].cpu().numpy(), np.zeros(2048, np.int64), want_adds=False, n_threads=O.max_threads())[0]
        out[f"add_n{n}"] = {"Mposes_s": m / t / 1e6, "bits_ok": bool(np.array_equal(add.view(np.uint32), ref.view(np.uint32)))}
    import hashlib
    for npts, Bn in ((500, 1 << 20), (1000, 1 << 18), (2048, 1 << 16)):
        pts = {9: W.box_mesh(npts, (0.1, 0.12, 0.05), 200 + npts)}
        tb_ = core.MeshTable(pts, {9: 0.1646}, pkg.SYMMETRIC_OBJECT_IDS, dev)
        ob = torch.full((Bn,), 9, dtype=torch.int64, device=dev)
        t = timed(lambda: tb_.evaluate(qb[:Bn], tb[:Bn], qa[:Bn], ta[:Bn], ob, want_adds=True), 5)
        packed = tb_.evaluate(qb[:Bn], tb[:Bn], qa[:Bn], ta[:Bn], ob, want_adds=True)[4]
        out[f"adds_n{npts}"] = {"Mposes_s": Bn / t / 1e6, "frac_nominal": Bn * 8 * npts * npts / t / 74.45e12,
                                "md5": hashlib.md5(packed[:10 * Bn].cpu().numpy().tobytes()).hexdigest()[:8],
                                "sched": tb_.schedule_state()}
    n = 1 << 22
    pq, gq, pt, gt = rnd(n, 4), rnd(n, 4), rnd(n, 3), rnd(n, 3)
    o3 = torch.empty(3, device=dev); g1 = torch.empty_like(pq); g2 = torch.empty_like(pt)
    ws = torch.zeros(64, dtype=torch.uint8, device=dev)
    t = timed(lambda: core.check(L.p6d_pose_loss_fwd_bwd(pq.data_ptr(), pt.data_ptr(), gq.data_ptr(), gt.data_ptr(), n, 1.0, 10.0, 0,
                                                         o3.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), 0, st)))
    out["pose_loss_4M_us"] = t * 1e6
    out["pose_loss_GBs"] = n * 84 / t / 1e9
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for lib in sys.argv[1:]:
            subprocess.run([sys.executable, __file__, "--one", lib])
