#!/usr/bin/env python
"""Time the ADD-S kernel variants (dev tool): python tools/variants.py [n_variants] [poses] [n_points]"""
import importlib, os, subprocess, sys, json
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

def one(poses, npts):
    import torch
    pkg = importlib.import_module("6d-pose-estimation_b200")
    W = pkg.workloads
    dev = torch.device("cuda", 0)
    pts, dia = W.config2_meshes(npts)
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, dev)
    d = [torch.from_numpy(x).to(dev) for x in W.config2(poses)]
    order = torch.argsort(d[4], stable=True).to(torch.int32)
    for _ in range(2):
        out = table.evaluate(*d, want_adds=True, order=order)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = table.evaluate(*d, want_adds=True, order=order); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    packed = out[4].cpu().numpy()
    import hashlib
    print(json.dumps({"ms": min(ts), "mposes": poses / min(ts) / 1e3, "hash": hashlib.md5(packed[:10 * poses].tobytes()).hexdigest()}))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), int(sys.argv[3]))
    else:
        # first argument: number of variants (0..n-1) or a comma list, "auto" = library's own choice
        arg = sys.argv[1] if len(sys.argv) > 1 else "8"
        vs = [x for x in arg.split(",")] if ("," in arg or arg == "auto") else [str(i) for i in range(int(arg))]
        poses = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
        npts = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
        for v in vs:
            env = dict(os.environ)
            env.pop("P6D_ADDS_VARIANT", None)
            if v != "auto":
                env["P6D_ADDS_VARIANT"] = v
            r = subprocess.run([sys.executable, __file__, "--one", str(poses), str(npts)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(v, line, flush=True)
