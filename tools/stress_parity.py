#!/usr/bin/env python
"""Large randomized GPU-vs-oracle parity run (not part of pytest: minutes of CPU time).
   python tools/stress_parity.py [poses_per_mesh_size]
Compares per-pose ADD, ADD-S (bit patterns), ADD-0.1d hit and validity for many mesh sizes,
pose noise levels, non-unit quaternions and unknown ids; prints one JSON line."""
import importlib, json, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("6d-pose-estimation_b200")
import oracle as O
W = pkg.workloads
dev = torch.device("cuda", 0)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
sizes = [5, 11, 64, 200, 500, 777, 1000, 1500, 2048, 3000]
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
tot = {"poses": 0, "add_mismatch": 0, "adds_mismatch": 0, "hit_mismatch": 0, "valid_mismatch": 0, "hits": 0,
       "add_only_mismatch": 0, "add_only_single_mesh_mismatch": 0, "add_only_poses": 0}
t0 = time.time()
for si, n in enumerate(sizes):
    B = max(256, int(per * min(1.0, (500.0 / n) ** 2)))       # keep the CPU side bounded for large meshes
    pts = {0: W.sphere_mesh(n, 0.12, 900 + si), 9: W.box_mesh(n, (0.1, 0.12, 0.05), 950 + si)}
    dia = {0: 0.12, 9: 0.1646}
    r = np.random.RandomState(1000 + si)
    pq, pt, gq, gt = W.random_poses(B, 2000 + si, rot_sigma=np.exp(r.uniform(np.log(1e-3), np.log(0.5), B)),
                                    trans_sigma=0.02 * r.rand(B, 1) ** 2)
    pq[::17] *= r.uniform(0.5, 1.5, (len(pq[::17]), 1)).astype(np.float32)        # non-unit quaternions
    obj = np.where(r.rand(B) < 0.5, 0, 9).astype(np.int64)
    obj[::101] = 4                                                                   # id without a mesh
    crit = pkg.ADDLoss(os.path.join(REPO, "tests", "golden"), dev)
    for k, v in pts.items():
        crit.points[k] = torch.from_numpy(v).to(dev)
    crit.diameters.update(dia)
    got = crit.eval_poses(*(torch.from_numpy(x).to(dev) for x in (pq, pt, gq, gt, obj)))
    ref = O.add_eval(O.MeshTable(pts, dia), pq, pt, gq, gt, obj, n_threads=O.max_threads())
    tot["poses"] += B
    tot["add_mismatch"] += int((bits(got["add"]) != bits(ref[0])).sum())
    tot["adds_mismatch"] += int((bits(got["add_s"]) != bits(ref[1])).sum())
    tot["hit_mismatch"] += int((got["hit"] != ref[2]).sum())
    tot["valid_mismatch"] += int((got["valid"] != ref[3]).sum())
    tot["hits"] += int(ref[2].sum())
    # kernel (a) on the same poses: two-mesh table (negotiated staging) and a single-mesh table (barrier-free form)
    dv = [torch.from_numpy(x).to(dev) for x in (pq, pt, gq, gt, obj)]
    a2 = crit._mesh_table(dev).evaluate(*dv, want_adds=False)[0].cpu().numpy()
    tot["add_only_mismatch"] += int((bits(a2) != bits(ref[0])).sum())
    one = pkg.core.MeshTable({0: pts[0]}, {0: dia[0]}, pkg.SYMMETRIC_OBJECT_IDS, dev)
    a1 = one.evaluate(*dv, want_adds=False)[0].cpu().numpy()
    sel = obj == 0
    tot["add_only_single_mesh_mismatch"] += int((bits(a1[sel]) != bits(ref[0][sel])).sum()) + int((a1[~sel] != 0).sum())
    tot["add_only_poses"] += 2 * B
tot["mesh_sizes"] = sizes
tot["seconds"] = round(time.time() - t0, 1)
print(json.dumps(tot))
