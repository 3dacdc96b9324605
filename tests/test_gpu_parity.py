"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes -> libp6d.so), against
the CPU oracle on the same seeded inputs and against the committed golden vectors that
came from the reference itself.

Bars (BASELINE.json north-star): ADD-0.1d decisions and per-object accuracy bit-exact;
distances and loss within 1e-5 relative.  Distances are in fact checked for bit equality
(the kernels reproduce the reference's float32 rounding and summation order); the
transcendental loss is checked at 1e-5.
"""
import json
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import bits, golden_meshes, load_golden, same_bits

pytestmark = pytest.mark.gpu

EVAL_CASES = ["eval_cfg1", "eval_cfg1_mixed", "eval_cfg2_subset", "eval_ragged", "eval_default_diameter",
              "eval_degenerate"]


def make_crit(pkg, pts, dia, dev):
    crit = pkg.ADDLoss(tempfile.mkdtemp(), dev)
    for k, v in pts.items():
        crit.points[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dev)
    crit.diameters.update(dia)
    return crit


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def test_native_library_is_loaded(pkg, cuda_dev):
    info = pkg.core.device_info(cuda_dev.index)
    assert info["cc"][0] >= 10, f"sm_100a cubin only; device is sm_{info['cc'][0]}{info['cc'][1]}"
    assert any("libp6d.so" in l for l in open("/proc/self/maps"))


def test_quat_to_mat(pkg, cuda_dev):
    g = load_golden("quat_to_mat")
    crit = pkg.ADDLoss(tempfile.mkdtemp(), cuda_dev)
    R = crit._quat_to_mat(T(g["q"], cuda_dev))
    assert R.shape == (64, 3, 3) and R.device == cuda_dev
    assert same_bits(R.cpu().numpy(), g["R"])


@pytest.mark.parametrize("name", EVAL_CASES)
def test_eval_matches_reference_golden(pkg, cuda_dev, name):
    g = load_golden(name)
    pts, dia = golden_meshes(g)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    args = [T(g[k], cuda_dev) for k in ("pq", "pt", "gq", "gt", "obj")]
    pp = crit.eval_poses(*args)
    assert np.array_equal(pp["valid"], g["valid"])
    assert np.array_equal(pp["hit"], g["hit"])            # bit-exact accept/reject
    assert same_bits(pp["add"], g["add"])                 # bit-exact distances (bar: 1e-5 rel)
    assert same_bits(pp["add_s"], g["adds"])
    m = crit.eval_metrics(*args)
    got = np.array([m["add_mean"], m["add_s_mean"], m["add_01d_acc"]], np.float64)
    assert np.array_equal(got, g["agg"], equal_nan=True)
    assert all(isinstance(v, np.float64) for v in m.values())


def test_eval_empty_and_unknown_ids(pkg, cuda_dev, W):
    crit = make_crit(pkg, {2: W.sphere_mesh(64, 0.1, 41)}, {}, cuda_dev)
    pq, pt, gq, gt = (T(x, cuda_dev) for x in W.random_poses(4, 42))
    m = crit.eval_metrics(pq, pt, gq, gt, torch.tensor([5, 5, 7, -1], device=cuda_dev))
    assert m == {"add_mean": 0, "add_s_mean": 0, "add_01d_acc": 0}
    assert all(isinstance(v, int) for v in m.values())
    m = crit.eval_metrics(pq[:0], pt[:0], gq[:0], gt[:0], torch.zeros(0, dtype=torch.long, device=cuda_dev))
    assert m == {"add_mean": 0, "add_s_mean": 0, "add_01d_acc": 0}
    empty = pkg.ADDLoss(tempfile.mkdtemp(), cuda_dev)
    assert empty.eval_metrics(pq, pt, gq, gt, torch.zeros(4, dtype=torch.long, device=cuda_dev))["add_mean"] == 0


def test_host_tensors_are_accepted_and_mutation_rebuilds_table(pkg, cuda_dev, W, oracle):
    pts = {0: W.sphere_mesh(300, 0.1, 7)}
    crit = make_crit(pkg, pts, {0: 0.1}, cuda_dev)
    pq, pt, gq, gt = W.random_poses(16, 8)
    obj = np.zeros(16, np.int64)
    a = crit.eval_poses(*(torch.from_numpy(x) for x in (pq, pt, gq, gt, obj)))     # CPU tensors in
    ref = oracle.add_eval(oracle.MeshTable(pts, {0: 0.1}), pq, pt, gq, gt, obj)
    assert same_bits(a["add"], ref[0]) and same_bits(a["add_s"], ref[1]) and np.array_equal(a["hit"], ref[2])
    # mutate the public dicts (tests and SURVEY 0.5 do this): the device table must follow
    pts2 = {0: W.sphere_mesh(180, 0.1, 9), 3: W.sphere_mesh(50, 0.2, 10)}
    crit.points[0] = T(pts2[0], cuda_dev); crit.points[3] = T(pts2[3], cuda_dev); crit.diameters[3] = 0.05
    obj[::2] = 3
    b = crit.eval_poses(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
    ref = oracle.add_eval(oracle.MeshTable(pts2, {0: 0.1, 3: 0.05}), pq, pt, gq, gt, obj)
    assert same_bits(b["add"], ref[0]) and same_bits(b["add_s"], ref[1]) and np.array_equal(b["hit"], ref[2])


def test_eval_large_mixed_batch_vs_oracle(pkg, cuda_dev, W, oracle):
    """2,600 poses over the 13 LineMOD ids (500-point meshes, the reference's default size),
    unsorted ids -> exercises the sorted-order path, mesh re-staging and the gt split."""
    pts, dia = W.sweep_meshes(500)
    r = np.random.RandomState(12)
    B = 2600
    pq, pt, gq, gt = W.random_poses(B, 13, rot_sigma=np.geomspace(0.005, 0.3, B))
    obj = np.array(W.LINEMOD_IDS, np.int64)[r.randint(0, 13, B)]
    obj[::97] = 2                                    # id without a mesh
    crit = make_crit(pkg, pts, dia, cuda_dev)
    got = crit.eval_poses(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq, pt, gq, gt, obj, n_threads=oracle.max_threads())
    assert np.array_equal(got["valid"], ref[3]) and np.array_equal(got["hit"], ref[2])
    assert same_bits(got["add"], ref[0]) and same_bits(got["add_s"], ref[1])
    assert 0 < got["hit"].sum() < got["valid"].sum()      # both decision outcomes exercised


def test_eval_ragged_large_meshes_vs_oracle(pkg, cuda_dev, W, oracle):
    """Mesh sizes around every switch of the ADD-S kernel: gt-split factor S (powers of two
    of threads*K), padded tails, one vs several passes over the pred points (N > 2048),
    cascade levels of the ordered mean (N >= 512, 8192)."""
    sizes = [128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2049, 3000, 4100, 5555]
    pts, dia = {}, {}
    for k, n in enumerate(sizes):
        oid = k if k < 9 else k + 2                      # keep 9/10 free, then use them as symmetric ids
        pts[oid] = W.sphere_mesh(n, 0.15, 700 + k); dia[oid] = 0.15
    pts[9] = W.box_mesh(1500, (0.1, 0.12, 0.05), 790); dia[9] = 0.1646
    pts[10] = W.box_mesh(2500, (0.04, 0.17, 0.04), 791); dia[10] = 0.1759
    ids = np.array(sorted(pts), np.int64)
    obj = np.repeat(ids, 2)
    pq, pt, gq, gt = W.random_poses(len(obj), 71, rot_sigma=0.03, trans_sigma=0.004)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    got = crit.eval_poses(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq, pt, gq, gt, obj, n_threads=oracle.max_threads())
    assert same_bits(got["add"], ref[0]) and same_bits(got["add_s"], ref[1])
    assert np.array_equal(got["hit"], ref[2]) and got["valid"].all()


def test_add_only_kernel_vs_oracle(pkg, cuda_dev, W, oracle):
    """Kernel (a): warp-per-pose ADD without the all-pairs part (packed f32x2 arithmetic, mesh
    staged by TMA in the row-pair layout).  Unsorted mixed objects (several staging passes per
    round), sorted order, unknown ids, NaN / identical poses (the square root's slow path)."""
    pts, dia = W.sweep_meshes(500)
    pts[1] = W.sphere_mesh(1000, dia[1], 77)
    pts[4] = W.sphere_mesh(37, dia[4], 78)
    B = 3000
    pq, pt, gq, gt = W.random_poses(B, 14, rot_sigma=np.geomspace(0.005, 0.3, B))
    obj = np.array(W.LINEMOD_IDS, np.int64)[np.random.RandomState(15).randint(0, 13, B)]
    obj[7] = 2; obj[8] = 99; obj[9] = -1              # ids without a mesh
    pq[20] = gq[20]; pt[20] = gt[20]                  # distance exactly 0
    pt[21, 0] = np.nan
    pq[22] *= 3.0
    core = pkg.core
    table = core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq, pt, gq, gt, obj, want_adds=False)
    order = torch.argsort(d[4], stable=True).to(torch.int32)
    for o in (None, order):
        add, adds, hit, valid, _ = table.evaluate(*d, want_adds=False, order=o)
        assert adds is None
        assert same_bits(add.cpu().numpy(), ref[0])
        assert np.array_equal(hit.cpu().numpy(), ref[2]) and np.array_equal(valid.cpu().numpy(), ref[3])


def test_add_only_kernel_every_row_shape(pkg, cuda_dev, W, oracle):
    """Kernel (a) over mesh sizes that exercise every branch of the ordered mean: fewer than 8
    points, odd / even numbers of 32-element steps, left-over vectors, scalar tails, the cascade
    levels (n >= 512), and all three torch.mm rounding modes (n = 1, 2..10, >= 11)."""
    sizes = [1, 2, 3, 7, 8, 10, 11, 31, 32, 33, 63, 64, 65, 95, 96, 100, 127, 128, 129, 255, 256, 500, 511, 512,
             513, 640, 1000, 1023, 1024, 1025, 2047, 2048, 3000, 4097]
    for lo in range(0, len(sizes), 12):
        chunk = sizes[lo:lo + 12]
        pts = {k: W.sphere_mesh(n, 0.1, 300 + n) for k, n in enumerate(chunk)}
        dia = {k: 0.1 for k in pts}
        B = 16 * len(chunk)
        pq, pt, gq, gt = W.random_poses(B, 31 + lo, rot_sigma=np.geomspace(0.005, 0.3, B))
        obj = (np.arange(B) % len(chunk)).astype(np.int64)
        table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
        d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
        add, _, hit, valid, _ = table.evaluate(*d, want_adds=False)
        ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq, pt, gq, gt, obj, want_adds=False)
        got = add.cpu().numpy()
        bad = np.nonzero(bits(got) != bits(ref[0]))[0]
        assert bad.size == 0, [(chunk[obj[i]], got[i], ref[0][i]) for i in bad[:5]]
        assert np.array_equal(hit.cpu().numpy(), ref[2]) and valid.cpu().numpy().all()


@pytest.mark.parametrize("n", [5, 96, 500, 1000, 1001, 2048])
def test_add_only_kernel_single_mesh_table(pkg, cuda_dev, W, oracle, n):
    """Kernel (a) with a table that holds ONE mesh takes the barrier-free form (the mesh is staged once
    per CTA and the warps never meet again): same bits, poses with another id are skipped, exact zeros
    (the repaired branch of the four-way square root), NaN and a ragged last round."""
    pts = {3: W.sphere_mesh(n, 0.12, 900 + n)}
    dia = {3: 0.12}
    B = 2053
    pq, pt, gq, gt = W.random_poses(B, 50 + n, rot_sigma=np.geomspace(0.005, 0.3, B))
    obj = np.full(B, 3, np.int64)
    obj[5] = 2; obj[6] = 77; obj[7] = -4
    pq[20] = gq[20]; pt[20] = gt[20]                  # every distance exactly 0
    pt[21, 2] = np.nan
    pq[22] *= 1e-3                                    # tiny rotation entries: denormal-range squares for some points
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq, pt, gq, gt, obj, want_adds=False)
    order = torch.randperm(B, generator=torch.Generator().manual_seed(n)).to(torch.int32).to(cuda_dev)
    for o in (None, order):
        add, _, hit, valid, _ = table.evaluate(*d, want_adds=False, order=o)
        assert same_bits(add.cpu().numpy(), ref[0])
        assert np.array_equal(hit.cpu().numpy(), ref[2]) and np.array_equal(valid.cpu().numpy(), ref[3])
    assert valid.cpu().numpy().sum() == B - 3


@pytest.mark.parametrize("single", [False, True])
def test_add_only_kernel_large_launch_batches_of_poses_per_warp(pkg, cuda_dev, W, oracle, single):
    """Kernel (a) hands a warp BATCHES of up to 32 poses once a launch is large enough (lane l prepares pose l,
    the matrices go through shared memory, batches are claimed through an atomic counter).  A 1.3 M-pose launch
    (32 poses per batch) must give the bits of the same poses in 4,096-pose launches (one pose per warp, the
    form the other tests pin to the oracle) -- unsorted (every batch mixes objects: one staging pass per
    object and round), sorted, with ids that have no mesh, exact zeros (the scalar re-evaluation), NaN -- and
    a random subset is compared with the oracle directly."""
    if single:
        pts = {3: W.sphere_mesh(500, 0.12, 901)}
        ids = np.array([3, 3, 3, 3, 3, 3, 3, 7], np.int64)
    else:
        pts = {1: W.sphere_mesh(500, 0.1, 911), 2: W.sphere_mesh(1000, 0.1, 912), 4: W.sphere_mesh(37, 0.1, 913),
               5: W.sphere_mesh(2048, 0.1, 914), 6: W.sphere_mesh(131, 0.1, 915), 9: W.sphere_mesh(8, 0.1, 916),
               10: W.sphere_mesh(1, 0.1, 917), 12: W.sphere_mesh(416, 0.1, 918)}
        ids = np.array([1, 2, 4, 5, 6, 9, 10, 12, 3, 99, -1, 1, 1, 2], np.int64)
    dia = {k: 0.1 for k in pts}
    B = 1_300_000
    pq, pt, gq, gt = W.random_poses(B, 61, rot_sigma=np.geomspace(0.005, 0.3, B))
    r = np.random.RandomState(62)
    obj = ids[r.randint(0, len(ids), B)]
    same = r.choice(B, 300, replace=False)
    pq[same] = gq[same]; pt[same] = gt[same]          # every distance exactly 0: the pose is re-evaluated by the scalar path
    pt[r.choice(B, 50, replace=False), 1] = np.nan
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
    small = [torch.cat([table.evaluate(*(x[lo:lo + 4096] for x in d), want_adds=False)[k] for lo in range(0, B, 4096)])
             for k in (0, 2, 3)]
    order = torch.argsort(d[4], stable=True).to(torch.int32)
    for o in (None, order):
        add, _, hit, valid, _ = table.evaluate(*d, want_adds=False, order=o)
        assert torch.equal(add.view(torch.int32), small[0].view(torch.int32))
        assert torch.equal(hit, small[1]) and torch.equal(valid, small[2])
    sel = np.concatenate([r.choice(B, 6000, replace=False), same[:40]])
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq[sel], pt[sel], gq[sel], gt[sel], obj[sel], want_adds=False,
                          n_threads=oracle.max_threads())
    sel_t = torch.from_numpy(sel).to(cuda_dev)
    assert same_bits(add[sel_t].cpu().numpy(), ref[0])
    assert np.array_equal(hit[sel_t].cpu().numpy(), ref[2]) and np.array_equal(valid[sel_t].cpu().numpy(), ref[3])
    assert int(valid.sum()) == int(np.isin(obj, list(pts)).sum())


def test_packed_sqrt_equals_sqrt_rn_on_every_float(pkg, cuda_dev):
    """sqrt2_rn and the unchecked four-way form + its range test (kernel (a)'s packed square roots) against
    sqrt.rn.f32 over all 2^32 bit patterns."""
    assert pkg.core.selftest_sqrt2(cuda_dev.index) == 0


def test_host_entry_and_accumulators(pkg, cuda_dev, W):
    """p6d_add_eval_host (host buffers, copies inside) == device entry; the per-object
    atomics equal the per-pose outputs summed on the host."""
    pts, dia = W.sweep_meshes(500)
    B = 1500
    pq, pt, gq, gt = W.random_poses(B, 21, rot_sigma=np.geomspace(0.01, 0.2, B))
    obj = np.array(W.LINEMOD_IDS, np.int64)[np.arange(B) % 13]
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    h = table.evaluate_host(pq, pt, gq, gt, obj)
    add, adds, hit, valid, _ = table.evaluate(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
    assert same_bits(h["add"], add.cpu().numpy()) and same_bits(h["adds"], adds.cpu().numpy())
    assert np.array_equal(h["hit"], hit.cpu().numpy()) and np.array_equal(h["valid"], valid.cpu().numpy())
    assert h["gpu_launches"] == 1
    for o in W.LINEMOD_IDS:
        sel = obj == o
        assert h["obj_valid"][o] == sel.sum() and h["obj_hits"][o] == h["hit"][sel].sum()
        assert np.isclose(h["obj_add_sum"][o], h["add"][sel].astype(np.float64).sum(), rtol=1e-12)
        assert np.isclose(h["obj_adds_sum"][o], h["adds"][sel].astype(np.float64).sum(), rtol=1e-12)


def test_config2_full_size_properties(pkg, cuda_dev, W, oracle):
    """BASELINE config 2 at full size (65,536 poses, 2,048-point meshes): size-independent
    properties plus a seeded subset against the oracle."""
    pts, dia = W.config2_meshes(2048)
    pq, pt, gq, gt, obj = W.config2(65536)
    # plant exact-match poses: every distance must be exactly 0 and the pose accepted
    pq[100:104] = gq[100:104]; pt[100:104] = gt[100:104]
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
    acc = [torch.zeros(table.n_slots, dtype=torch.int64, device=cuda_dev) for _ in range(2)] + \
          [torch.zeros(table.n_slots, dtype=torch.float64, device=cuda_dev) for _ in range(2)]
    order = torch.argsort(d[4], stable=True).to(torch.int32)
    add, adds, hit, valid, _ = table.evaluate(*d, want_adds=True, order=order, acc=acc)
    add, adds, hit, valid = (x.cpu().numpy() for x in (add, adds, hit, valid))
    assert valid.all()
    assert np.all(adds <= add)                       # min over j includes j == i, same rounding
    assert not np.any(add[100:104]) and not np.any(adds[100:104]) and hit[100:104].all()
    # decisions are consistent with the float64 threshold compare on the returned distances
    thr = np.where(obj == 9, 0.1 * dia[9], 0.1 * dia[10])
    assert np.array_equal(hit.astype(bool), adds.astype(np.float64) < thr)
    # per-object accumulators (the quantities that get all-reduced) equal the per-pose sums
    for o in (9, 10):
        assert int(acc[0][o]) == int(hit[obj == o].sum()) and int(acc[1][o]) == int((obj == o).sum())
    # processing order does not change any bit
    perm = np.random.RandomState(3).permutation(65536)[:8192]
    add2, adds2, hit2, _, _ = table.evaluate(*(x[torch.from_numpy(perm).to(cuda_dev)].contiguous() for x in d))
    assert same_bits(add2.cpu().numpy(), add[perm]) and same_bits(adds2.cpu().numpy(), adds[perm])
    assert np.array_equal(hit2.cpu().numpy(), hit[perm])
    # seeded subset against the oracle (bit-exact)
    sel = np.random.RandomState(4).choice(65536, 96, replace=False)
    ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq[sel], pt[sel], gq[sel], gt[sel], obj[sel],
                          n_threads=oracle.max_threads())
    assert same_bits(add[sel], ref[0]) and same_bits(adds[sel], ref[1]) and np.array_equal(hit[sel], ref[2])


def test_mesh_too_large_is_a_loud_error(pkg, cuda_dev, W):
    n_max = pkg.core.C.c_int()
    pkg.core.check(pkg.core.lib().p6d_adds_max_points(cuda_dev.index, pkg.core.C.byref(n_max)))
    assert n_max.value >= 4096
    crit = make_crit(pkg, {0: W.sphere_mesh(n_max.value + 64, 0.1, 1)}, {0: 0.1}, cuda_dev)
    pq, pt, gq, gt = (T(x, cuda_dev) for x in W.random_poses(2, 2))
    with pytest.raises(pkg.core.P6DError, match="at most"):
        crit.eval_metrics(pq, pt, gq, gt, torch.zeros(2, dtype=torch.long, device=cuda_dev))
    with pytest.raises(ValueError):
        crit.eval_metrics(pq, pt[:1], gq, gt, torch.zeros(2, dtype=torch.long, device=cuda_dev))


def test_batch_beyond_the_launch_limit_is_refused(pkg, cuda_dev, W):
    # the pose scheduler claims work through a 32-bit counter: the C ABI refuses larger launches
    crit = make_crit(pkg, {0: W.sphere_mesh(64, 0.1, 1)}, {0: 0.1}, cuda_dev)
    table = crit._mesh_table(cuda_dev)
    buf = torch.zeros(64, dtype=torch.float32, device=cuda_dev)
    obj = torch.zeros(8, dtype=torch.long, device=cuda_dev)
    out = torch.zeros(64, dtype=torch.uint8, device=cuda_dev)
    L, ptr = pkg.core.lib(), pkg.core.ptr
    rc = L.p6d_add_eval(table.handle, ptr(buf), ptr(buf), ptr(buf), ptr(buf), ptr(obj), None, 2 ** 31,
                        ptr(buf), ptr(buf), ptr(out), ptr(out), None, None, None)
    assert rc == pkg.core.P6D_EINVAL and b"split the batch" in L.p6d_last_error()
    torch.cuda.synchronize(cuda_dev)


def test_add_forward_value(pkg, cuda_dev):
    g = load_golden("add_forward")
    pts, dia = golden_meshes(g)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    args = [T(g[k], cuda_dev) for k in ("pq", "pt", "gq", "gt", "obj")]
    loss = crit(*args)
    assert loss.dim() == 0 and loss.device == cuda_dev
    # one launch, the reference's grouping on the device: the very bits of the reference's loss
    assert bits(np.float32(loss.item())) == bits(g["loss"])
    assert bits(np.float32(crit.train_loss(*args).item())) == bits(g["loss"])
    none = crit(args[0][:2], args[1][:2], args[2][:2], args[3][:2], torch.tensor([6, 6], device=cuda_dev))
    assert none.item() == 0.0 and none.requires_grad
    none.backward()                                   # must not raise (reference: fresh requires_grad leaf)
    with torch.no_grad():
        assert not crit(*args).requires_grad


@pytest.mark.parametrize("name", ["add_forward_n500", "add_forward_small", "add_forward_mixed"])
def test_add_forward_value_more_goldens(pkg, cuda_dev, name):
    """Loss form against further reference outputs: the reference's default 500-point meshes with a
    32-sample batch; meshes of <= 44 points (ATen's naive bmm kernel rounds differently from MKL);
    a batch mixing both with unknown ids.  Bit equality."""
    g = load_golden(name)
    pts, dia = golden_meshes(g)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    args = [T(g[k], cuda_dev) for k in ("pq", "pt", "gq", "gt", "obj")]
    assert bits(np.float32(crit(*args).item())) == bits(g["loss"])


def test_add_forward_backward(pkg, cuda_dev, W, oracle):
    """N3: gradient of ADDLoss.forward vs the reference's autograd (golden) and vs the oracle
    at a larger size (N = 2048 symmetric + asymmetric, where the reference needs ~GBs)."""
    g = load_golden("add_forward")
    pts, dia = golden_meshes(g)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    pq = T(g["pq"], cuda_dev).requires_grad_(True); pt = T(g["pt"], cuda_dev).requires_grad_(True)
    loss = crit(pq, pt, T(g["gq"], cuda_dev), T(g["gt"], cuda_dev), T(g["obj"], cuda_dev))
    assert loss.requires_grad
    (2.0 * loss).backward()
    for got, ref in ((pq.grad, g["grad_q"]), (pt.grad, g["grad_t"])):
        ref = 2.0 * ref
        scale = np.maximum(np.abs(ref).max(1, keepdims=True), 1e-30)
        assert np.all(np.abs(got.cpu().numpy() - ref) <= 1e-5 * scale)       # 1e-5 relative per row
    pts2 = {3: W.sphere_mesh(2048, 0.17, 5), 10: W.box_mesh(2048, (0.04, 0.17, 0.04), 6)}
    dia2 = {3: 0.17, 10: 0.176}
    crit2 = make_crit(pkg, pts2, dia2, cuda_dev)
    a, b, c, d = W.random_poses(6, 77, rot_sigma=0.1, trans_sigma=0.01)
    obj = np.array([3, 10, 10, 3, 8, 10], np.int64)
    x = T(a, cuda_dev).requires_grad_(True); y = T(b, cuda_dev).requires_grad_(True)
    crit2(x, y, T(c, cuda_dev), T(d, cuda_dev), T(obj, cuda_dev)).backward()
    rq, rt = oracle.add_backward(oracle.MeshTable(pts2, dia2), a, b, c, d, obj)
    for got, ref in ((x.grad, rq), (y.grad, rt)):
        scale = np.maximum(np.abs(ref).max(1, keepdims=True), 1e-30)
        assert np.all(np.abs(got.cpu().numpy() - ref) <= 1e-5 * scale)
    assert not torch.any(x.grad[4])
    # B >= 256 takes the sorted-order path of the forward kernel; value and grads must not care
    pts3, dia3 = W.sweep_meshes(120)
    crit3 = make_crit(pkg, pts3, dia3, cuda_dev)
    B = 600
    a, b, c, d = W.random_poses(B, 78, rot_sigma=0.05, trans_sigma=0.01)
    obj = np.array(W.LINEMOD_IDS, np.int64)[np.random.RandomState(79).randint(0, 13, B)]
    x = T(a, cuda_dev).requires_grad_(True); y = T(b, cuda_dev).requires_grad_(True)
    loss = crit3(x, y, T(c, cuda_dev), T(d, cuda_dev), T(obj, cuda_dev))
    loss.backward()
    ref = oracle.add_forward(oracle.MeshTable(pts3, dia3), a, b, c, d, obj)
    assert abs(loss.item() - float(ref)) <= 1e-5 * float(ref)
    rq, rt = oracle.add_backward(oracle.MeshTable(pts3, dia3), a, b, c, d, obj)
    for got, r in ((x.grad, rq), (y.grad, rt)):
        scale = np.maximum(np.abs(r).max(1, keepdims=True), 1e-30)
        assert np.all(np.abs(got.cpu().numpy() - r) <= 1e-5 * scale)


# ------------------------------------------------------------------ PoseLoss
@pytest.mark.parametrize("mode", ["geodesic", "l1"])
@pytest.mark.parametrize("tag,B", [("b32", 32), ("b5", 5)])
def test_pose_loss_forward_backward_golden(pkg, cuda_dev, mode, tag, B):
    g = load_golden("pose_loss_cfg3")
    k = f"{mode}_{tag}_"
    rot = T(g[k + "rot_in"], cuda_dev).requires_grad_(True)
    trans = T(g[k + "trans_in"], cuda_dev).requires_grad_(True)
    crit = pkg.PoseLoss(1.0, 10.0, mode)
    loss = crit(rot, trans, T(g["gt_rot"][:B], cuda_dev), T(g["gt_trans"][:B], cuda_dev))
    assert loss.dim() == 0
    loss.backward()
    assert abs(loss.item() - float(g[k + "loss"])) <= 1e-5 * abs(float(g[k + "loss"]))   # north-star tolerance
    gq = g[k + "grad_rot"]
    row_scale = np.maximum(np.abs(gq).max(1, keepdims=True), 1e-30)
    assert np.all(np.abs(rot.grad.cpu().numpy() - gq) <= 1e-5 * row_scale)
    assert same_bits(trans.grad.cpu().numpy(), g[k + "grad_trans"])
    fn = crit._geodesic_distance if mode == "geodesic" else crit._quaternion_l1
    rl = fn(T(g[k + "rot_in"], cuda_dev), T(g["gt_rot"][:B], cuda_dev)).item()
    assert abs(rl - float(g[k + "rot"])) <= 1e-5 * abs(float(g[k + "rot"]))
    if mode == "l1":   # no transcendental: the ordered mean makes it bit-exact
        assert bits(np.float32(rl)) == bits(g[k + "rot"])


def test_training_step_chain_rgb_geometric(pkg, cuda_dev):
    """Config 3 as the training script runs it (reference train_rgb_geometric.py:104-108):
    head output -> q/(|q|+1e-8) in autograd -> pinhole translation -> PoseLoss -> backward;
    grads reach rot_raw and z_pred."""
    g = load_golden("pose_loss_cfg3")
    rot_raw = T(g["rot_raw"], cuda_dev).requires_grad_(True)
    z = T(g["z_pred"], cuda_dev).requires_grad_(True)
    rot = rot_raw / (torch.norm(rot_raw, dim=1, keepdim=True) + 1e-8)
    trans = pkg.pinhole_translation(z, T(g["bbox_center"], cuda_dev), T(g["K"], cuda_dev))
    loss = pkg.PoseLoss(1.0, 10.0, "geodesic")(rot, trans, T(g["gt_rot"], cuda_dev), T(g["gt_trans"], cuda_dev))
    loss.backward()
    assert abs(loss.item() - float(g["geodesic_b32_loss"])) <= 1e-5 * float(g["geodesic_b32_loss"])
    assert np.allclose(z.grad.cpu().numpy(), g["geodesic_b32_grad_z"], rtol=1e-5, atol=1e-7)
    ref = g["geodesic_b32_grad_rot_raw"]
    got = rot_raw.grad.cpu().numpy()
    finite = np.isfinite(ref).all(1)
    scale = np.maximum(np.abs(ref[finite]).max(1, keepdims=True), 1e-30)
    assert np.all(np.abs(got[finite] - ref[finite]) <= 2e-5 * scale)


def test_fused_geometric_training_step(pkg, cuda_dev, W):
    """(d1) fused into (c): one launch == pinhole_translation + PoseLoss, bit for bit, for the
    loss, the translation and the gradients w.r.t. rot and z (small and multi-CTA batch)."""
    for B, mode in ((32, "geodesic"), (32, "l1"), (5000, "geodesic")):
        c = W.config3(B, 3)
        crit = pkg.PoseLoss(1.0, 10.0, mode)
        for Kt in (T(c["K"], cuda_dev), T(c["K"][0], cuda_dev)):
            r1 = T(c["rot_raw"], cuda_dev).requires_grad_(True); z1 = T(c["z_pred"], cuda_dev).requires_grad_(True)
            t1 = pkg.pinhole_translation(z1, T(c["bbox_center"], cuda_dev), Kt)
            l1 = crit(r1, t1, T(c["gt_rot"], cuda_dev), T(c["gt_trans"], cuda_dev))
            l1.backward()
            r2 = T(c["rot_raw"], cuda_dev).requires_grad_(True); z2 = T(c["z_pred"], cuda_dev).requires_grad_(True)
            l2, t2 = crit.forward_geometric(r2, z2, T(c["bbox_center"], cuda_dev), Kt, T(c["gt_rot"], cuda_dev),
                                            T(c["gt_trans"], cuda_dev))
            l2.backward()        # upstream gradient 1: every rounding happens in the same place
            assert same_bits(t2.cpu().numpy(), t1.detach().cpu().numpy())
            assert same_bits(r2.grad.cpu().numpy(), r1.grad.cpu().numpy())
            assert same_bits(z2.grad.cpu().numpy(), z1.grad.cpu().numpy()) and z2.grad.shape == (B, 1)
            if B <= 2048:
                assert l2.item() == l1.item()
            else:   # float64 atomics: summation order of the block partials is not fixed
                assert abs(l2.item() - l1.item()) <= 1e-6 * abs(l1.item())
    # non-unit upstream gradient: scaled once at the end instead of before the pinhole backward
    c = W.config3(32, 3)
    r = T(c["rot_raw"], cuda_dev).requires_grad_(True); z = T(c["z_pred"], cuda_dev).requires_grad_(True)
    l, _ = pkg.PoseLoss(1.0, 10.0).forward_geometric(r, z, T(c["bbox_center"], cuda_dev), T(c["K"], cuda_dev),
                                                     T(c["gt_rot"], cuda_dev), T(c["gt_trans"], cuda_dev))
    (3.0 * l).backward()
    g = load_golden("pose_loss_cfg3")
    assert np.allclose(z.grad.cpu().numpy(), 3.0 * g["geodesic_b32_grad_z"], rtol=1e-5, atol=1e-7)


def test_captured_training_step_equals_the_eager_step(pkg, cuda_dev, W):
    """PoseLoss.capture: loss + gradients of config 3 replayed from a CUDA graph equal the eager
    criterion(...).backward() bit for bit, for new inputs written into the static tensors, in both
    forms (plain and forward_geometric)."""
    crit = pkg.PoseLoss(1.0, 10.0, "geodesic")
    c = W.config3(32, 6)
    Tn = lambda x: T(x, cuda_dev)
    step = crit.capture(Tn(c["rot_raw"]), Tn(c["gt_trans"]) + 0.01, Tn(c["gt_rot"]), Tn(c["gt_trans"]))
    geo = crit.capture(Tn(c["rot_raw"]), Tn(c["z_pred"]), Tn(c["gt_rot"]), Tn(c["gt_trans"]),
                       geometric=(Tn(c["bbox_center"]), Tn(c["K"])))
    for seed in (6, 7, 8):
        c = W.config3(32, seed)
        rot = Tn(c["rot_raw"]).requires_grad_(True)
        tr = (Tn(c["gt_trans"]) + 0.01 * seed).requires_grad_(True)
        loss = crit(rot, tr, Tn(c["gt_rot"]), Tn(c["gt_trans"]))
        loss.backward()
        got = step(rot, tr, Tn(c["gt_rot"]), Tn(c["gt_trans"]))
        assert got.item() == loss.item()
        assert torch.equal(step.grad_rot, rot.grad) and torch.equal(step.grad_trans, tr.grad)
        # inputs the kernel cannot read in place (float64, a strided view) take the copy + replay path
        step.grad_rot.zero_(); step.grad_trans.zero_()
        wide = torch.stack([tr.detach(), tr.detach()], 1)[:, 0]
        assert not wide.is_contiguous()
        got = step(rot.detach().double(), wide, Tn(c["gt_rot"]), Tn(c["gt_trans"]))
        assert got.item() == loss.item()
        assert torch.equal(step.grad_rot, rot.grad) and torch.equal(step.grad_trans, tr.grad)
        rot2 = Tn(c["rot_raw"]).requires_grad_(True)
        z = Tn(c["z_pred"]).requires_grad_(True)
        l2, trans = crit.forward_geometric(rot2, z, Tn(c["bbox_center"]), Tn(c["K"]), Tn(c["gt_rot"]), Tn(c["gt_trans"]))
        l2.backward()
        g2 = geo(rot2, z, Tn(c["bbox_center"]), Tn(c["K"]), Tn(c["gt_rot"]), Tn(c["gt_trans"]))
        assert g2.item() == l2.item() and torch.equal(geo.translation, trans)
        assert torch.equal(geo.grad_rot, rot2.grad) and torch.equal(geo.grad_z, z.grad)


def test_pose_loss_weights_large_batch_and_no_grad(pkg, cuda_dev, W, oracle):
    g = load_golden("pose_loss_cfg3")
    a = T(g["rot_raw"], cuda_dev).requires_grad_(True)
    b = T(g["pred_trans_direct"], cuda_dev).requires_grad_(True)
    loss = pkg.PoseLoss(0.5, 2.0)(a, b, T(g["gt_rot"], cuda_dev), T(g["gt_trans"], cuda_dev), obj_ids=None)
    (3.0 * loss).backward()                                   # upstream gradient != 1
    assert abs(loss.item() - float(g["w_loss"])) <= 1e-5 * float(g["w_loss"])
    ref = 3.0 * g["w_grad_rot"]
    scale = np.maximum(np.abs(ref).max(1, keepdims=True), 1e-30)
    assert np.all(np.abs(a.grad.cpu().numpy() - ref) <= 1e-5 * scale)
    assert np.allclose(b.grad.cpu().numpy(), 3.0 * g["w_grad_trans"], rtol=1e-6)
    # multi-CTA shape (B > 2048) against the oracle
    B = 50000
    pq, pt, gq, gt = W.random_poses(B, 33, rot_sigma=0.3, trans_sigma=0.05)
    for mode in ("geodesic", "l1"):
        x = T(pq, cuda_dev).requires_grad_(True); y = T(pt, cuda_dev).requires_grad_(True)
        l = pkg.PoseLoss(1.0, 10.0, mode)(x, y, T(gq, cuda_dev), T(gt, cuda_dev))
        l.backward()
        o = oracle.pose_loss(pq, pt, gq, gt, 1.0, 10.0, mode)
        assert abs(l.item() - float(o["loss"])) <= 1e-5 * float(o["loss"])
        sc = np.maximum(np.abs(o["grad_q"]).max(1, keepdims=True), 1e-30)
        assert np.all(np.abs(x.grad.cpu().numpy() - o["grad_q"]) <= 1e-5 * sc)
        assert same_bits(y.grad.cpu().numpy(), o["grad_t"])
        l2 = pkg.PoseLoss(1.0, 10.0, mode)(x.detach(), y.detach(), T(gq, cuda_dev), T(gt, cuda_dev))
        assert not l2.requires_grad and l2.item() == l.item()      # workspace left clean, deterministic


def test_large_pose_loss_streamed_and_direct_kernels_agree_and_misalignment_is_refused(pkg, cuda_dev, W, oracle):
    """B > 2048: 16-byte aligned inputs go through the TMA-streamed kernel, anything else through the
    grid-stride kernel; both against the oracle.  Misaligned float4 rows are an error, not a fault."""
    core = pkg.core
    L = core.lib()
    B = 30000 + 77                                           # 117 full tiles + a ragged tail
    pq, pt, gq, gt = W.random_poses(B, 35, rot_sigma=0.3, trans_sigma=0.05)
    o = oracle.pose_loss(pq, pt, gq, gt, 1.0, 10.0, "geodesic")
    ws = torch.zeros(64, dtype=torch.uint8, device=cuda_dev)
    dq, dgq = T(pq, cuda_dev), T(gq, cuda_dev)
    pad = torch.zeros(3 * B + 8, dtype=torch.float32, device=cuda_dev)
    res = {}
    for name, off in (("streamed", 0), ("direct", 1)):      # off = 1 float: [B,3] rows only 4-byte aligned
        dpt = pad[off:off + 3 * B].copy_(T(pt, cuda_dev).reshape(-1))
        dgt = torch.zeros(3 * B + 8, dtype=torch.float32, device=cuda_dev)[off:off + 3 * B].copy_(T(gt, cuda_dev).reshape(-1))
        assert (dpt.data_ptr() % 16 == 0) == (off == 0)
        out = torch.empty(3, device=cuda_dev); g1 = torch.empty(B, 4, device=cuda_dev); g2 = torch.empty(3 * B, device=cuda_dev)
        core.check(L.p6d_pose_loss_fwd_bwd(dq.data_ptr(), dpt.data_ptr(), dgq.data_ptr(), dgt.data_ptr(), B, 1.0, 10.0, 0,
                                           out.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), cuda_dev.index,
                                           core.stream_ptr(cuda_dev)))
        torch.cuda.synchronize()
        res[name] = (out.cpu().numpy(), g1.cpu().numpy(), g2.cpu().numpy().reshape(B, 3))
        assert abs(res[name][0][0] - float(o["loss"])) <= 1e-5 * float(o["loss"]), name
        sc = np.maximum(np.abs(o["grad_q"]).max(1, keepdims=True), 1e-30)
        assert np.all(np.abs(res[name][1] - o["grad_q"]) <= 1e-5 * sc), name
        assert same_bits(res[name][2], o["grad_t"]), name
    assert same_bits(res["streamed"][1], res["direct"][1])      # same row arithmetic in both kernels
    rc = L.p6d_pose_loss_fwd_bwd(dq.data_ptr() + 4, pad.data_ptr(), dgq.data_ptr(), pad.data_ptr(), 16, 1.0, 10.0, 0,
                                 ws.data_ptr(), None, None, ws.data_ptr(), cuda_dev.index, core.stream_ptr(cuda_dev))
    assert rc == core.P6D_EINVAL and b"16-byte aligned" in L.p6d_last_error()


# ------------------------------------------------------------------ geometric translation
def test_pinhole_forward_backward(pkg, cuda_dev):
    g = load_golden("pinhole")
    z = T(g["z"], cuda_dev).requires_grad_(True)
    out = pkg.pinhole_translation(z, T(g["uv"], cuda_dev), T(g["K"], cuda_dev))
    assert same_bits(out.detach().cpu().numpy(), g["out"])
    out.backward(T(g["grad_out"], cuda_dev))
    assert z.grad.shape == (32, 1)
    assert np.allclose(z.grad.cpu().numpy(), g["grad_z"], rtol=1e-5, atol=1e-7)
    out = pkg.pinhole_translation(T(g["z"], cuda_dev), T(g["uv"], cuda_dev), T(g["K_shared"], cuda_dev))
    assert same_bits(out.cpu().numpy(), g["out_shared"])


def test_pinhole_shared_k_vector_kernel_equals_the_row_kernel(pkg, cuda_dev, oracle):
    """Shared [3,3] K and B >= 1024: four rows per thread with 16-byte accesses; B % 4 = 1, 2, 3 tails."""
    r = np.random.RandomState(9)
    K = pkg.DEFAULT_K.astype(np.float32)
    for B in (1024, 4097, 10002, 65539):
        z = (r.rand(B, 1) * 1.2 + 0.3).astype(np.float32)
        uv = (r.rand(B, 2) * np.array([640, 480])).astype(np.float32)
        got = pkg.pinhole_translation(T(z, cuda_dev), T(uv, cuda_dev), T(K, cuda_dev)).cpu().numpy()
        ref, _ = oracle.pinhole(z, uv, K)
        assert same_bits(got, ref), B


def test_depth_backproject(pkg, cuda_dev, W, oracle):
    g = load_golden("depth_backproject")
    depth, uv, K = W.config4(256, int(g["seed"]))
    out = pkg.depth_backproject(T(depth, cuda_dev), T(uv, cuda_dev), T(K, cuda_dev))
    assert same_bits(out.cpu().numpy(), g["out"])
    out = pkg.depth_backproject(T(depth, cuda_dev), T(uv, cuda_dev), T(K[0], cuda_dev))
    assert same_bits(out.cpu().numpy(), g["out_shared"])
    # other crop size / clamp (explicit H, W instead of the hard-coded 224)
    depth, uv, K = W.config4(64, 44, hw=(96, 128), clamp_hi=95)
    out = pkg.depth_backproject(T(depth, cuda_dev), T(uv, cuda_dev), T(K, cuda_dev), clamp_hi=95.0)
    assert same_bits(out.cpu().numpy(), oracle.depth_backproject(depth, uv, K, clamp_hi=95.0))
    with pytest.raises(pkg.core.P6DError):
        pkg.depth_backproject(T(depth, cuda_dev), T(uv, cuda_dev), T(K, cuda_dev), clamp_hi=223.0)


# ------------------------------------------------------------------ sweep (N2 / multi-GPU sharding)
def test_sweep_is_invariant_to_the_number_of_ranks_and_matches_oracle(pkg, cuda_dev, W, oracle):
    """Config-5-shaped sweep at reduced size through the native runner (p6d_sweep_run): counts are
    identical for 1 rank and for the sum of 2 / 3 emulated ranks (same GPU, no process group); the
    first poses of every block, as evaluated, agree with the oracle bit for bit; a block regenerated
    through the public generator + translation kernels reproduces what the sweep evaluated."""
    pts = {0: W.sphere_mesh(200, 0.102, 1), 9: W.box_mesh(160, (0.1, 0.12, 0.05), 2), 12: W.sphere_mesh(90, 0.278, 3)}
    dia = {0: 0.102, 9: 0.1646, 12: 0.278}
    n, chunk = 3000, 1024
    full, launches, check = pkg.evaluate_sweep(pts, dia, cuda_dev, n, chunk=chunk, seed=77, check_n=64)
    # per block 3 chunks x (generate + evaluate + chunk totals) and one translation launch per chunk of the two
    # geometric variants
    assert launches == 3 * 3 * (4 * 3 + 2) and int(full.valid.sum()) == 3 * 4 * n
    for world in (2, 3):
        parts = [pkg.evaluate_sweep(pts, dia, cuda_dev, n, chunk=chunk, seed=77, rank=r, world=world)[0]
                 for r in range(world)]
        assert torch.equal(sum(p.hits for p in parts), full.hits)          # integer-exact for any G
        assert torch.equal(sum(p.valid for p in parts), full.valid)
        assert torch.allclose(sum(p.add_sum for p in parts), full.add_sum, rtol=1e-12)
    tab = full.table(pkg.sweep.VARIANTS)
    assert set(tab) == set(pkg.sweep.VARIANTS) and tab["rgb"][9]["n"] == n
    # every block's first 64 hypotheses against the oracle
    table = oracle.MeshTable(pts, dia)
    nb = check["pq"].shape[0]
    assert nb == 12
    flat = lambda k: check[k].reshape(nb * 64, -1).squeeze()
    ref = oracle.add_eval(table, flat("pq"), flat("pt"), flat("gq"), flat("gt"), check["obj"].reshape(-1),
                          n_threads=oracle.max_threads())
    assert same_bits(flat("add"), ref[0]) and same_bits(flat("adds"), ref[1])
    assert np.array_equal(flat("hit"), ref[2])
    # block (object index 1 = id 9): regenerate each variant through the public pieces
    K = torch.tensor(pkg.DEFAULT_K, dtype=torch.float32, device=cuda_dev)
    for vi, var in enumerate(pkg.sweep.VARIANTS):
        blk = pkg.sweep.synth_block(64, 0, 77, 1, vi, var, 9, cuda_dev)
        pt = pkg.sweep.variant_translation(var, blk, K)
        row = 1 * 4 + vi
        assert same_bits(blk["pq"].cpu().numpy(), check["pq"][row]) and same_bits(blk["gt"].cpu().numpy(), check["gt"][row])
        assert same_bits(pt.cpu().numpy(), check["pt"][row])
        if var.endswith("geometric"):      # the geometric heads land near the GT translation
            assert torch.allclose(pt[:, 2], blk["gt"][:, 2], atol=0.03) and torch.allclose(pt[:, :2], blk["gt"][:, :2], atol=0.01)
    # the generator is a pure function of the hypothesis index: a shifted window overlaps exactly
    a1 = pkg.sweep.synth_block(100, 0, 77, 2, 3, "rgbd_geometric", 12, cuda_dev)
    a2 = pkg.sweep.synth_block(60, 40, 77, 2, 3, "rgbd_geometric", 12, cuda_dev)
    for k in ("pq", "gq", "gt", "uv", "kc", "depth"):
        assert torch.equal(a1[k][40:], a2[k]), k
    q = a1["gq"]
    assert torch.allclose(q.norm(dim=1), torch.ones(100, device=cuda_dev), atol=1e-5)


def test_depth_crop_backproject_fused(pkg, cuda_dev, W, oracle):
    """N1 on the GPU: 4 texels per box instead of a padded crop + 224x224 resize."""
    g = load_golden("crop_backproject")
    depth, boxes = W.config4_frame(int(g["seed"]), 256)
    K = torch.from_numpy(g["K"]).to(cuda_dev)
    for mode, tag in (("cv2", "optimized"), ("generic", "generic")):     # == reference dataset + model, bit for bit
        xyz, center, kcrop, zmm = pkg.depth_crop_backproject(torch.from_numpy(depth).to(cuda_dev), torch.from_numpy(boxes),
                                                             K, return_aux=True, bilinear=mode)
        assert same_bits(center.cpu().numpy(), g[f"{tag}_center"]) and same_bits(kcrop.cpu().numpy(), g[f"{tag}_Kcrop"])
        assert np.array_equal(zmm.cpu().numpy(), g[f"{tag}_z_mm"])
        assert same_bits(xyz.cpu().numpy(), g[f"{tag}_xyz"])
    # the fixture whose pixels separate cv2's default (IPP) arithmetic from its generic one
    gi = load_golden("crop_backproject_ipp")
    frame, _ = W.config4_frame(int(gi["seed"]), 8)
    for mode in ("cv2", "generic"):
        xyz, center, kcrop, zmm = pkg.depth_crop_backproject(torch.from_numpy(frame).to(cuda_dev),
                                                             torch.from_numpy(gi["boxes"]), K, return_aux=True, bilinear=mode)
        assert np.array_equal(zmm.cpu().numpy(), gi[f"{mode}_z_mm"]), mode
        assert same_bits(xyz.cpu().numpy(), gi[f"{mode}_xyz"]) and same_bits(center.cpu().numpy(), gi[f"{mode}_center"])
    assert int((gi["cv2_z_mm"] != gi["generic_z_mm"]).sum()) >= 64
    # the fused result equals the two-step API (crop tensors + p6d_depth_backproject) as well
    # other seeds / frame sizes against the oracle, boxes partly outside the frame
    depth2, boxes2 = W.config4_frame(41, 300, hw=(360, 500))
    boxes2[:8, 0] -= 60; boxes2[8:16, 1] += 200
    for mode in ("cv2", "generic"):
        r = oracle.crop_depth_backproject(depth2, boxes2, g["K"], bilinear=mode)
        xyz2 = pkg.depth_crop_backproject(depth2, boxes2, K, bilinear=mode)
        assert same_bits(xyz2.cpu().numpy(), r["xyz"]), mode
    assert pkg.depth_crop_backproject(depth2, boxes2[:0], K).shape == (0, 3)


def test_detection_backproject_inference_form(pkg, cuda_dev, W, oracle):
    """N1, inference form on the GPU (integer xyxy detector boxes, float32 resize, float64 centre / K_crop):
    bit-equal to the values the reference's own inference-script lines + model method produced, and to the
    oracle on another frame size with boxes partly outside the frame."""
    g = load_golden("crop_backproject_xyxy")
    depth, _ = W.config4_frame(int(g["seed"]), 256)
    xyz, center, kcrop, zm = pkg.detection_backproject(torch.from_numpy(depth).to(cuda_dev), g["boxes_xyxy"], g["K"],
                                                       return_aux=True)
    assert same_bits(center.cpu().numpy(), g["center"]) and same_bits(kcrop.cpu().numpy(), g["Kcrop"])
    assert same_bits(zm.cpu().numpy(), g["z_m"]) and same_bits(xyz.cpu().numpy(), g["xyz"])
    # default K = DEFAULT_K (float64), as in the script
    assert same_bits(pkg.detection_backproject(torch.from_numpy(depth).to(cuda_dev), g["boxes_xyxy"]).cpu().numpy(), g["xyz"])
    depth2, b2 = W.config4_frame(43, 3000, hw=(360, 500))
    xyxy = np.stack([b2[:, 0], b2[:, 1], b2[:, 0] + b2[:, 2], b2[:, 1] + b2[:, 3]], 1).astype(np.int32)
    xyxy[:8, [0, 2]] -= 60; xyxy[8:16, [1, 3]] += 200; xyxy[16] = (5, 5, 5, 5)      # the last one: empty box -> fallback
    r = oracle.detection_depth_backproject(depth2, xyxy[:16], g["K"])
    got = pkg.detection_backproject(torch.from_numpy(depth2).to(cuda_dev), xyxy, g["K"], return_aux=True)
    assert same_bits(got[0][:16].cpu().numpy(), r["xyz"]) and same_bits(got[3][:16].cpu().numpy(), r["z_m"])
    assert got[0][16].tolist() == [0.0, 0.0, 0.5]
    r = oracle.detection_depth_backproject(depth2, xyxy[17:600], g["K"])
    assert same_bits(got[0][17:600].cpu().numpy(), r["xyz"]) and same_bits(got[1][17:600].cpu().numpy(), r["center"])
    assert pkg.detection_backproject(torch.from_numpy(depth2).to(cuda_dev), xyxy[:0]).shape == (0, 3)


def _edge_poses(W, B, seed, sigma):
    pq, pt, gq, gt = W.random_poses(B, seed, rot_sigma=sigma)
    pq[3] = gq[3]; pt[3] = gt[3]                      # identical poses: every distance 0
    pt[4, 0] = np.nan                                 # NaN propagates through every minimum
    pq[5] = np.nan
    pt[6, 2] = np.inf
    pq[7] *= 3.0                                      # non-unit quaternion: R is not a rotation (spheres must not assume it)
    pq[8] = 0.0                                       # zero quaternion: R = I scaled to ... whatever _quat_to_mat gives
    gq[9] *= 1e-3
    return pq, pt, gq, gt


@pytest.mark.parametrize("sizes", [(2048, 1000), (3000, 33, 64), (500, 131, 1), (777, 1024, 1500, 2)])
def test_exact_pruned_adds_kernel_equals_all_pairs(pkg, cuda_dev, W, oracle, sizes):
    """The opt-in pruned ADD-S kernel (b') must return every byte the all-pairs kernel returns -- ADD, ADD-S,
    decisions, valid flags, borderline flags -- for good and bad predictions, mixed objects in any order, ids
    without a mesh, NaN / inf / zero / non-unit poses; a subset is compared with the oracle directly."""
    pts = {k + 1: (W.sphere_mesh(n, 0.1, 400 + n) if k % 2 == 0 else W.box_mesh(n, (0.1, 0.12, 0.05), 500 + n))
           for k, n in enumerate(sizes)}
    dia = {k: 0.1 for k in pts}
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    B = 6000
    for sigma in (0.01, 0.1, 1.0):
        pq, pt, gq, gt = _edge_poses(W, B, 70 + len(sizes), np.full(B, sigma))
        obj = np.array(list(pts) + [0, 99, -3], np.int64)[np.random.RandomState(71).randint(0, len(pts) + 3, B)]
        d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
        order = torch.argsort(d[4], stable=True).to(torch.int32)
        for o in (None, order):
            full = table.evaluate(*d, order=o)[4]
            prun = table.evaluate(*d, order=o, prune=True)[4]
            assert torch.equal(full[:11 * B], prun[:11 * B]), (sizes, sigma)
        sel = np.random.RandomState(72).choice(B, 64, replace=False)
        ref = oracle.add_eval(oracle.MeshTable(pts, dia), pq[sel], pt[sel], gq[sel], gt[sel], obj[sel],
                              n_threads=oracle.max_threads())
        adds = prun[4 * B:8 * B].view(torch.float32).cpu().numpy()
        assert same_bits(adds[sel], ref[1])


def test_exact_pruning_switch_degenerate_meshes_and_limits(pkg, cuda_dev, W, oracle):
    """The per-table switch (ADDLoss.exact_pruning / MeshTable.set_pruning) changes no result; meshes whose
    blocks degenerate (all points equal, collinear points, a NaN vertex) stay exact; a mesh beyond the pruned
    kernel's shared-memory limit is refused by the explicit entry and silently takes the all-pairs kernel
    under the switch."""
    r = np.random.RandomState(80)
    line = np.zeros((900, 3), np.float32); line[:, 0] = np.linspace(-0.05, 0.05, 900)
    same = np.full((800, 3), 0.01, np.float32)
    bad = W.sphere_mesh(1000, 0.1, 81).copy(); bad[17, 1] = np.nan
    pts = {1: line, 2: same, 3: bad, 4: W.sphere_mesh(1200, 0.1, 82)}
    dia = {k: 0.1 for k in pts}
    B = 2000
    pq, pt, gq, gt = _edge_poses(W, B, 83, np.geomspace(0.005, 0.5, B))
    obj = r.randint(1, 5, B).astype(np.int64)
    d = [T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)]
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    full = table.evaluate(*d)[4]
    assert torch.equal(full[:11 * B], table.evaluate(*d, prune=True)[4][:11 * B])
    table.set_pruning(True)
    assert torch.equal(full[:11 * B], table.evaluate(*d)[4][:11 * B])          # through p6d_add_eval, switch on
    host = table.evaluate_host(pq, pt, gq, gt, obj)                              # and through the host entry
    assert same_bits(host["adds"], full[4 * B:8 * B].view(torch.float32).cpu().numpy())
    # ADDLoss surface: the aggregate dict does not change with the switch
    crit = make_crit(pkg, {k: v for k, v in pts.items() if k != 3}, dia, cuda_dev)
    keep = obj != 3
    args = [T(x[keep], cuda_dev) for x in (pq, pt, gq, gt, obj)]
    m0 = crit.eval_metrics(*args)
    crit.exact_pruning = True
    m1 = crit.eval_metrics(*args)
    assert {k: repr(v) for k, v in m0.items()} == {k: repr(v) for k, v in m1.items()}
    # beyond the pruned kernel's limit
    big = pkg.core.MeshTable({0: W.sphere_mesh(6000, 0.1, 84)}, {0: 0.1}, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    e = [T(x[:8], cuda_dev) for x in (pq, pt, gq, gt)] + [torch.zeros(8, dtype=torch.long, device=cuda_dev)]
    with pytest.raises(pkg.core.P6DError, match="at most"):
        big.evaluate(*e, prune=True)
    ref = big.evaluate(*e)[4]
    big.set_pruning(True)
    assert torch.equal(ref[:88], big.evaluate(*e)[4][:88])


def test_two_tables_of_different_size_interleaved(pkg, cuda_dev, W, oracle):
    """The ADD-S kernel's shared-memory attribute is per device, not per table: a small and a
    large table used alternately must both keep launching (and keep their results)."""
    small = ({0: W.sphere_mesh(1200, 0.1, 1)}, {0: 0.1})
    big = ({0: W.sphere_mesh(4000, 0.1, 2)}, {0: 0.1})
    pq, pt, gq, gt = W.random_poses(6, 3)
    obj = np.zeros(6, np.int64)
    crits = [make_crit(pkg, *small, cuda_dev), make_crit(pkg, *big, cuda_dev)]
    refs = [oracle.add_eval(oracle.MeshTable(*m), pq, pt, gq, gt, obj, n_threads=4) for m in (small, big)]
    for _ in range(2):
        for crit, ref in zip(crits, refs):
            got = crit.eval_poses(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
            assert same_bits(got["add_s"], ref[1]) and same_bits(got["add"], ref[0])


def test_concurrent_launches_on_one_table_from_two_threads(pkg, cuda_dev, W, oracle):
    """include/p6d.h: p6d_add_eval may be called from several host threads on one table; each
    launch takes its own scheduler counter, so launches overlapping on two streams keep their bits."""
    import threading
    mesh = ({9: W.sphere_mesh(700, 0.12, 5)}, {9: 0.12})
    crit = make_crit(pkg, *mesh, cuda_dev)
    table = crit._mesh_table(cuda_dev)
    batches = []
    for seed in (11, 12):
        pq, pt, gq, gt = W.random_poses(3000, seed)
        obj = np.full(3000, 9, np.int64)
        ref = oracle.add_eval(oracle.MeshTable(*mesh), pq[:64], pt[:64], gq[:64], gt[:64], obj[:64], n_threads=4)
        batches.append(([T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)], ref))
    torch.cuda.synchronize(cuda_dev)
    results, errors = [None, None], []

    def worker(i):
        try:
            stream = torch.cuda.Stream(device=cuda_dev)
            with torch.cuda.stream(stream):
                outs = [table.evaluate(*batches[i][0]) for _ in range(8)]
            stream.synchronize()
            results[i] = [(o[0].cpu().numpy(), o[1].cpu().numpy(), o[2].cpu().numpy()) for o in outs]
        except Exception as e:  # surfaced below
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for i in range(2):
        ref = batches[i][1]
        first = results[i][0]
        assert same_bits(first[0][:64], ref[0]) and same_bits(first[1][:64], ref[1])
        assert np.array_equal(first[2][:64], ref[2])
        for other in results[i][1:]:
            assert all(np.array_equal(a, b) for a, b in zip(first, other))


@pytest.mark.parametrize("n_points", [500, 900, 1500, 2048])
def test_relaid_scan_loop_equals_the_ptxas_schedule(pkg, cuda_dev, W, n_points):
    """The ADD-S scan loops re-laid after linking (csrc/sass_sched.py) against the ptxas-scheduled
    kernels of the same mesh-size class: every output byte over 16,384 seeded poses, through the
    library's own self-check entry (the check it runs on 592 poses before first use on a device)."""
    pts, dia = W.config2_meshes(n_points)
    table = pkg.core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, cuda_dev)
    assert table.selfcheck(16384) == 0
    st = table.schedule_state()
    assert st["runtime_state"] in ((1,) if st["built_relaid"] else (0,))
    # a first launch leaves the class verified (or the build has nothing re-laid); never rejected
    d = [T(x, cuda_dev) for x in W.config2(64)]
    table.evaluate(*d)
    assert table.schedule_state()["runtime_state"] != 2


def test_project_points_batch(pkg, cuda_dev):
    """N4 on the GPU: float64 batched projection == the reference's per-pose NumPy result."""
    import importlib
    u = importlib.import_module("6d-pose-estimation_b200.utils")
    g = load_golden("projection")
    uv = u.project_points_batch(g["corners"], g["quat"], g["trans"], g["K"], device=cuda_dev)
    assert uv.shape == (64, 8, 2) and uv.dtype == torch.int64
    assert np.array_equal(uv.cpu().numpy(), g["uv"])
    uv = u.project_points_batch(g["corners"], g["Rmat"], g["trans"], g["K"], device=cuda_dev)
    assert np.array_equal(uv.cpu().numpy(), g["uv"])


def test_gpu_invariants_sign_and_identity(pkg, cuda_dev, W):
    """GPU twin of tests/test_properties.py: q and -q give the same bits; pred == gt gives
    exact zeros and acceptance; ADD-S <= ADD."""
    pts, dia = W.sweep_meshes(500)
    crit = make_crit(pkg, pts, dia, cuda_dev)
    B = 1300
    pq, pt, gq, gt = W.random_poses(B, 91, rot_sigma=np.geomspace(1e-3, 0.4, B))
    obj = np.array(W.LINEMOD_IDS, np.int64)[np.arange(B) % 13]
    a = crit.eval_poses(*(T(x, cuda_dev) for x in (pq, pt, gq, gt, obj)))
    b = crit.eval_poses(*(T(x, cuda_dev) for x in (-pq, pt, -gq, gt, obj)))
    assert same_bits(a["add"], b["add"]) and same_bits(a["add_s"], b["add_s"]) and np.array_equal(a["hit"], b["hit"])
    assert np.all(a["add_s"] <= a["add"])
    z = crit.eval_poses(*(T(x, cuda_dev) for x in (gq, gt, gq, gt, obj)))
    assert not z["add"].any() and not z["add_s"].any() and z["hit"].all()
