#!/usr/bin/env python
"""sched_search.py -- dev tool (runs on a GPU box): hill-climb the schedule of one re-laid scan loop
with live timings.  Starts from the recipe the Makefile applies (csrc/sass_sched.py), then tries
single changes -- flip one yield bit, move one minimum by one slot -- and keeps a change when the
kernel gets faster.  Every candidate goes through the same safety checks as the build pass and its
output hash is compared with the baseline's.

  python tools/sched_search.py UNPATCHED.so KERNEL VARIANT N_POINTS POSES BUDGET_SECONDS [policy] [yield]
Prints the best schedule as --order / --yield-mask arguments for sass_sched.py.
"""
import ctypes as C
import hashlib, importlib, json, os, random, struct, sys, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "6d-pose-estimation_b200", "csrc"))
import sass_sched as S
import numpy as np
import torch

base_so, kernel, variant, npts, poses, budget = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
policy = sys.argv[7] if len(sys.argv) > 7 else "spaced=FADD2:2"
yspec = sys.argv[8] if len(sys.argv) > 8 else "8,0"
os.environ["P6D_ADDS_VARIANT"] = variant

pkg = importlib.import_module("6d-pose-estimation_b200")
core, W = pkg.core, pkg.workloads
dev = torch.device("cuda", 0)
pts, dia = W.config2_meshes(npts)
d = [torch.from_numpy(x).to(dev) for x in W.config2(poses)]
order_t = torch.argsort(d[4], stable=True).to(torch.int32)

kernel_instrs = S.load(base_so, kernel)
body = S.pick_loop(kernel_instrs, "uniform")
data0 = bytearray(open(base_so, "rb").read())
whole = b"".join(struct.pack("<QQ", i.lo, i.hi) for i in kernel_instrs)
assert data0.count(whole) == 1
loop_at = data0.find(whole) + (body[0].addr - kernel_instrs[0].addr)
mark_at = data0.find(S.STATE_OLD)
n = len(body)
counter = [0]


def evaluate(order, yields, packed_stall=1):
    S.check_order(body, order)
    stalls, _ = S.assign_stalls(body, order, packed_stall)
    blob = S.emit(body, order, stalls, yields)
    data = bytearray(data0)
    data[loop_at:loop_at + len(blob)] = blob
    if mark_at >= 0:
        data[mark_at:mark_at + len(S.STATE_NEW)] = S.STATE_NEW
    counter[0] += 1
    path = f"/dev/shm/p6d_cand_{os.getpid()}_{counter[0]}.so"
    open(path, "wb").write(data)
    core._lib = None
    core.SO_PATH = path
    table = core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, dev)
    out = table.evaluate(*d, want_adds=True, order=order_t)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = table.evaluate(*d, want_adds=True, order=order_t); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    h = hashlib.md5(out[4].cpu().numpy()[:10 * poses].tobytes()).hexdigest()
    table.close()
    os.unlink(path)
    return min(ts), h


if policy.endswith(".json"):          # continue from an earlier result
    plan = json.load(open(policy))
    order, yields = plan["order"], [int(c) for c in plan["yield_mask"]]
else:
    order = S.make_order(body, policy)
    per, ph = (int(x) for x in yspec.split(","))
    yields = [0 if (q % per) == ph else 1 for q in range(n)]
ident = list(range(n))
t_ptxas, h0 = evaluate(ident, [i.field()["y"] for i in body], 2)
best_t, h = evaluate(order, yields)
assert h == h0
print(f"ptxas schedule {t_ptxas:.3f} ms; start recipe {best_t:.3f} ms", flush=True)
t_end = time.time() + budget
rng = random.Random(int(os.environ.get('SEARCH_SEED', '1')))
evals, accepted = 0, 0
def mutate(cand_o, cand_y):
    if rng.random() < 0.5:
        p = rng.randrange(n - 1)
        cand_y[p] ^= 1
        return f"yield[{p}]"
    mins = [p for p, k in enumerate(cand_o) if S.movable(body[k])]
    p = rng.choice(mins)
    q = p + rng.choice((-1, 1))
    if q < 0 or q >= n - 1:
        return None
    cand_o[p], cand_o[q] = cand_o[q], cand_o[p]
    cand_y[p], cand_y[q] = cand_y[q], cand_y[p]
    return f"min {p}->{q}"


while time.time() < t_end:
    cand_o, cand_y = list(order), list(yields)
    # single moves have converged once; 40 % of the candidates now combine two or three changes
    k_moves = 1 if rng.random() < 0.6 else rng.choice((2, 3))
    what = [mutate(cand_o, cand_y) for _ in range(k_moves)]
    if None in what:
        continue
    what = "+".join(what)
    try:
        t, h = evaluate(cand_o, cand_y)
    except (SystemExit, AssertionError):
        continue
    evals += 1
    if h != h0:
        print("HASH MISMATCH for", what, flush=True)
        continue
    if t < best_t - 0.015:
        best_t, order, yields = t, cand_o, cand_y
        accepted += 1
        print(f"{evals:4d} {what:28s} -> {best_t:.3f} ms", flush=True)
print(json.dumps({"kernel": kernel, "evals": evals, "accepted": accepted, "ptxas_ms": t_ptxas, "best_ms": best_t,
                  "order": order, "yield_mask": "".join(str(y) for y in yields)}))
