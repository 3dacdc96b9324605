#!/usr/bin/env python
"""Load balance of the persistent ADD-S grid (dev tool): python tools/timeline.py [poses]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
W = pkg.workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
pts, dia = W.config2_meshes(2048)
table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, dev)
d = [torch.from_numpy(x).to(dev) for x in W.config2(B)]
order = torch.argsort(d[4], stable=True).to(torch.int32)
for rep in range(2):
    tl = table.timeline(*d, order=order).astype(np.int64)
t0 = tl[:, 1].min()
dur = (tl[:, 2] - tl[:, 1]) / 1e6
print(f"ctas={len(tl)} span={(tl[:,2].max()-t0)/1e6:.3f} ms  start spread={(tl[:,1].max()-t0)/1e6:.3f} ms")
print(f"cta duration ms: min={dur.min():.3f} p50={np.median(dur):.3f} max={dur.max():.3f}")
print(f"end time ms: min={(tl[:,2].min()-t0)/1e6:.3f} max={(tl[:,2].max()-t0)/1e6:.3f}")
print(f"poses per cta: min={tl[:,3].min()} p50={int(np.median(tl[:,3]))} max={tl[:,3].max()} sum={tl[:,3].sum()}")
sm = {}
for smid, s, e, n in tl:
    sm.setdefault(int(smid), []).append(int(n))
per_sm = np.array([sum(v) for v in sm.values()])
print(f"sms={len(sm)} ctas/sm={sorted(set(len(v) for v in sm.values()))} poses per sm: min={per_sm.min()} max={per_sm.max()}")
pairs = [sorted(v) for v in sm.values() if len(v) == 2]
if pairs:
    pa = np.array(pairs)
    print(f"within-SM split (slow cta, fast cta): mean={pa.mean(0)}, max ratio={np.max(pa[:,1]/np.maximum(pa[:,0],1)):.2f}")
