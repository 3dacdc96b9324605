"""CPU: host-side logic of the package and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden, same_bits


def test_library_exports_every_declared_symbol(pkg):
    """libp6d.so loads without a GPU and exports every function include/p6d.h declares (the
    P6D_DEV section belongs to the development build, `make -C csrc dev`, and is not shipped)."""
    header = open(os.path.join(REPO, "include", "p6d.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"#ifdef P6D_DEV.*?#endif", "", header, flags=re.S)
    declared = set(re.findall(r"\b(p6d_[a-z0-9_]+)\s*\(", header))
    assert {"p6d_add_eval", "p6d_add_eval_host", "p6d_pose_loss_fwd_bwd", "p6d_pinhole_fwd", "p6d_add_forward",
            "p6d_depth_backproject", "p6d_mesh_table_create", "p6d_sweep_run", "p6d_adds_selfcheck"} <= declared
    L = ctypes.CDLL(pkg.core.SO_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(pkg.core.EXPORTS) == declared
    assert pkg.core.lib().p6d_version() == pkg.core.P6D_VERSION == 4
    # the product library carries none of the development hooks
    assert not hasattr(L, "p6d_adds_timeline")
    blob = open(pkg.core.SO_PATH, "rb").read()
    assert b"P6D_ADDS_VARIANT" not in blob and b"P6D_DEBUG_SCAN_REPS" not in blob


def test_measured_schedule_plans_fit_the_loops_this_source_compiles_to(pkg):
    """csrc/sched_plan_*.json are tied to the instruction order ptxas emits (loop fingerprint): a
    source change that makes ptxas emit another loop silently drops the build to the generic recipe
    (-2 ... 3 % throughput).  The build log says which schedule every class got; with the committed
    plans and this toolchain (nvcc 12.9) all three must be the measured plan.  A different toolchain
    is a legitimate reason to see "recipe" here: re-run tools/sched_search.py then."""
    from pathlib import Path
    log = Path(pkg.core.SO_PATH).parent / "libp6d.sched.log"
    if not log.exists():
        pytest.skip("library built without the scheduling log (NOSCHED=1 or an older Makefile)")
    got = dict(line.split() for line in log.read_text().splitlines() if line.strip())
    assert got == {"n512": "plan", "n1024": "plan", "n2048": "plan"}, got


def test_post_link_scheduling_pass_and_its_checks(pkg, tmp_path):
    """The build re-lays the scan loops of the ADD-S kernels (csrc/sass_sched.py) and marks each
    mesh-size class it re-laid; a class it refused keeps ptxas' schedule (both are valid builds: at
    run time the library compares a re-laid kernel with its ptxas twin before using it).  No GPU is
    needed to check the pass itself."""
    import json, shutil, subprocess, sys
    from pathlib import Path
    csrc = Path(pkg.core.SO_PATH).parent
    blob = open(pkg.core.SO_PATH, "rb").read()
    at = blob.find(b"P6D-SCHED-STATE:")
    assert at >= 0 and blob.count(b"P6D-SCHED-STATE:") == 1
    state = blob[at + 16:at + 19].decode()
    assert set(state) <= {"p", "t"} and len(state) == 3
    assert pkg.core.lib().p6d_adds_schedule() == (1 if "t" in state else 0)
    sched = [sys.executable, str(csrc / "sass_sched.py")]
    ptxas_kernel = "adds_cta_kernelILi512ELi4ELi2ELi2ELi0EE"      # never touched by the pass
    relaid = {0: "adds_cta_kernelILi128ELi4ELi8ELi0ELi0EE", 1: "adds_cta_kernelILi256ELi4ELi4ELi0ELi0EE",
              2: "adds_cta_kernelILi256ELi8ELi2ELi0ELi0EE"}
    # (1) a failing run leaves its output untouched; a loop that is already re-laid is refused
    work = tmp_path / "lib.so"
    shutil.copy(pkg.core.SO_PATH, work)
    before = work.read_bytes()
    for cls, kern in relaid.items():
        if state[cls] != "t":
            continue
        r = subprocess.run(sched + [str(work), kern, "spaced=FADD2:2", "--loop=uniform", "--packed-stall=1",
                                    f"--out={work}", f"--mark={cls}"], capture_output=True, text=True)
        assert r.returncode != 0 and ("already" in r.stdout + r.stderr), r.stdout + r.stderr
        assert work.read_bytes() == before and not Path(str(work) + ".sched-tmp").exists()
    # (2) a plan measured on another instruction order of the loop is not applied
    plan = json.load(open(csrc / "sched_plan_n2048.json"))
    plan["loop_fingerprint"] = "0" * 32
    (tmp_path / "plan.json").write_text(json.dumps(plan))
    r = subprocess.run(sched + [str(work), f"--plan={tmp_path / 'plan.json'}", "--loop=uniform", f"--out={work}"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "another instruction order" in r.stdout + r.stderr
    assert work.read_bytes() == before
    # (3) the committed plans are permutations with one yield bit per instruction and name their loop
    for name, n_instr in (("sched_plan_n2048.json", 118), ("sched_plan_n1024.json", 62), ("sched_plan_n512.json", 62)):
        plan = json.load(open(csrc / name))
        assert sorted(plan["order"]) == list(range(n_instr)) and len(plan["yield_mask"]) == n_instr
        assert plan["order"][-1] == n_instr - 1 and plan["best_ms"] < plan["ptxas_ms"]
        assert len(plan["loop_fingerprint"]) == 32
    sys.path.insert(0, str(csrc))
    try:
        import sass_sched as S
    finally:
        sys.path.pop(0)
    # (4) the ptxas-scheduled twin of the large class still carries the compiler's stalls
    loops = [body for _, body in S.loops(S.load(pkg.core.SO_PATH, ptxas_kernel)) if any(i.op in S.PACKED for i in body)]
    assert loops
    for body in loops:
        packed = [i for i in body if i.op in S.PACKED]
        assert 4 * sum(i.field()["stall"] < 2 for i in packed) <= len(packed) and len(S.fingerprint(body)) == 32
    # an order that lets a minimum cross the packed op that overwrites its operand is refused
    sys.path.insert(0, str(csrc))
    try:
        import sass_sched as S
    finally:
        sys.path.pop(0)
    # (5) on an unscheduled copy of a software-pipelined loop: the recipe, a broken order, random releases
    unsched = tmp_path / "unsched.so"
    r = subprocess.run(["make", "-C", str(csrc), "NOSCHED=1", f"TARGET={unsched.name}", f"NEW={unsched.name}.new"],
                       capture_output=True, text=True)
    built = csrc / unsched.name
    assert r.returncode == 0 and built.exists(), r.stdout[-2000:] + r.stderr[-2000:]
    shutil.move(str(built), unsched)
    assert b"P6D-SCHED-STATE:ppp" in unsched.read_bytes()
    body = S.pick_loop(S.load(str(unsched), relaid[0]), "uniform")
    assert sum(S.movable(i) for i in body) == 8 and sum(i.op in S.PACKED for i in body) == 48
    order = S.make_order(body, "spaced=FADD2:2")
    S.check_order(body, order)
    stalls, _ = S.assign_stalls(body, order, 1)
    assert all(1 <= x <= 15 for x in stalls)
    first_min = next(p for p, k in enumerate(order) if S.movable(body[k]))
    writer = next(p for p in range(first_min + 1, len(order))
                  if S.conflicts(body[order[first_min]], body[order[p]]))
    bad = list(order)
    bad.insert(writer, bad.pop(first_min))      # now behind the instruction it conflicts with
    with pytest.raises(SystemExit):
        S.check_order(body, bad)
    # whatever release pattern a policy chooses, `sink` only produces orders the checker accepts,
    # and the stall assignment keeps every modelled read-after-write gap
    import random
    rng = random.Random(3)
    for _ in range(25):
        o = S.sink(body, lambda b, held, cur, nxt: held[:1] if (held and rng.random() < 0.3) else [])
        S.check_order(body, o)
        enc, _ = S.assign_stalls(body, o, rng.choice((1, 2)))
        t = S.model_times(body, o, enc)
        last_write = {}
        for p_, k_ in enumerate(o):
            ins = body[k_]
            for r in ins.src:
                q = last_write.get(r)
                if q is not None and not body[o[q]].var_lat:
                    lat = S.latency(body[o[q]], ins)
                    assert lat is None or t[p_] - t[q] >= lat
            for r in ins.dst:
                last_write[r] = p_
    # (6) the full pass on the unscheduled copy: patched, marked, read back
    r = subprocess.run(sched + [str(unsched), relaid[0], "spaced=FADD2:2", "--loop=uniform", "--packed-stall=1",
                                "--yield=period8,0", f"--out={unsched}", "--mark=0"], capture_output=True, text=True)
    assert r.returncode == 0 and "patched" in r.stdout, r.stdout + r.stderr
    assert b"P6D-SCHED-STATE:tpp" in unsched.read_bytes()


def test_no_cpu_fallback(pkg):
    crit = pkg.ADDLoss(tempfile.mkdtemp(), "cpu")
    crit.points[0] = torch.zeros(16, 3)
    z = lambda k: torch.zeros(2, k)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit.eval_metrics(z(4), z(3), z(4), z(3), torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.PoseLoss()(z(4), z(3), z(4), z(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.pinhole_translation(torch.ones(2, 1), z(2), torch.eye(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.depth_backproject(torch.ones(2, 224, 224), z(2), torch.eye(3))


def test_backward_views_of_the_pose_loss_buffer(pkg):
    """PoseLoss.backward takes each gradient as ONE as_strided view of the scaled buffer: the strides it
    computes are torch's contiguous strides, and the views equal the slice + view they replaced."""
    rm = pkg.models.pose_loss._row_major
    for shape in [(32, 4), (32, 3), (32, 1), (32,), (5, 2, 4), (1, 4), (7, 1, 3)]:
        assert rm(torch.Size(shape)) == torch.empty(shape).stride()
    B = 6
    buf = torch.arange((7 * B + 3) // 4 * 4 + 4, dtype=torch.float32)
    scaled = buf * 3.0
    rs, ts = torch.Size((B, 4)), torch.Size((3, 2, 3))
    assert torch.equal(scaled.as_strided(rs, rm(rs), 0), (buf[:7 * B] * 3.0)[:4 * B].view(rs))
    assert torch.equal(scaled.as_strided(ts, rm(ts), 4 * B), (buf[:7 * B] * 3.0)[4 * B:].view(ts))


def test_current_stream_accessor_has_a_fallback(pkg, monkeypatch):
    """stream_ptr uses torch's raw accessor when it exists and torch.cuda.current_stream otherwise."""
    core = pkg.core
    monkeypatch.setattr(core, "_raw_stream", lambda index: 1234 + index)
    assert core.stream_ptr(torch.device("cuda", 3)) == 1237
    calls = []

    class _S:
        cuda_stream = 77
    monkeypatch.setattr(core, "_raw_stream", None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: calls.append(d) or _S())
    assert core.stream_ptr(torch.device("cuda", 0)) == 77 and calls == [torch.device("cuda", 0)]


def test_product_never_imports_the_oracle():
    """Nothing under the package imports, loads or links oracle/ (comments may cite it)."""
    root = os.path.join(REPO, "6d-pose-estimation_b200")
    for d, _, files in os.walk(root):
        for f in files:
            path = os.path.join(d, f)
            if f.endswith(".py"):
                for line in open(path):
                    code = line.split("#")[0]
                    assert not re.match(r"\s*(import|from)\s+oracle\b", code), (f, line)
                    assert "libpose_oracle" not in code and "pose_oracle" not in code, (f, line)
            elif f.endswith((".cu", ".cuh", ".h")) or f == "Makefile":
                src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
                src = re.sub(r"//[^\n]*", "", src)
                src = re.sub(r"#[^\n]*oracle[^\n]*", "", src) if f == "Makefile" else src
                assert "p6o_" not in src and "pose_oracle" not in src and "oracle/" not in src, f


def test_surface_matches_reference_signatures(pkg):
    import inspect
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(pkg.ADDLoss.__init__) == ["self", "model_dir", "device", "rot_weight", "trans_weight"]
    assert sig(pkg.ADDLoss.eval_metrics) == ["self", "pred_r", "pred_t", "gt_r", "gt_t", "obj_ids"]
    assert sig(pkg.ADDLoss.forward) == ["self", "pred_r", "pred_t", "gt_r", "gt_t", "obj_ids"]
    assert sig(pkg.ADDLoss.train_loss) == sig(pkg.ADDLoss.forward)
    assert sig(pkg.ADDLoss._quat_to_mat) == ["self", "q"]
    assert sig(pkg.PoseLoss.__init__) == ["self", "rot_weight", "trans_weight", "rotation_loss"]
    d = inspect.signature(pkg.PoseLoss.__init__).parameters
    assert (d["rot_weight"].default, d["trans_weight"].default, d["rotation_loss"].default) == (1.0, 1.0, "geodesic")
    assert sig(pkg.PoseLoss.forward) == ["self", "pred_rot", "pred_trans", "gt_rot", "gt_trans", "obj_ids"]
    assert sig(pkg.get_gt_and_K) == ["data_dir", "obj_id_str", "frame_id"]
    assert pkg.SYMMETRIC_OBJECT_IDS == {9, 10}
    p = pkg.PoseLoss(2.0, 3.0, "l1")
    assert (p.rot_weight, p.trans_weight, p.rotation_loss_type) == (2.0, 3.0, "l1")
    # an empty batch gives NaN like the reference's means over zero rows (pose_loss.py:26,50), not an error
    z = lambda k: torch.zeros(0, k, requires_grad=True)
    out = pkg.PoseLoss()(z(4), z(3), torch.zeros(0, 4), torch.zeros(0, 3))
    assert out.dim() == 0 and torch.isnan(out) and out.requires_grad


def test_surface_matches_the_reference_itself(pkg):
    """With the reference reachable (build container: /root/reference, or P6D_REFERENCE), every public
    method / function of its three hot-path modules exists here with the same parameter names,
    order and defaults -- introspected, not hard-coded."""
    import importlib.util, inspect, sys
    ref = os.environ.get("P6D_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("reference not reachable on this machine (the hard-coded check above still runs)")

    def load(rel, name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    r_add, r_pose, r_cam = load("models/add_loss.py", "_ref_add_loss"), load("models/pose_loss.py", "_ref_pose_loss"), \
        load("utils/camera.py", "_ref_camera")
    ours = {"ADDLoss": pkg.ADDLoss, "PoseLoss": pkg.PoseLoss}

    def params(f):
        return [(n, p.default) for n, p in inspect.signature(f).parameters.items()]
    checked = 0
    for cls_name, ref_cls in (("ADDLoss", r_add.ADDLoss), ("PoseLoss", r_pose.PoseLoss)):
        for name, fn in vars(ref_cls).items():
            if not callable(fn) or name.startswith("__") and name != "__init__":
                continue
            mine = getattr(ours[cls_name], name, None)
            assert mine is not None, f"{cls_name}.{name} missing"
            got = params(mine)
            want = params(fn)
            if name == "_load_models":        # ours exposes the reference's hard-coded 500-point cap as a default
                got = got[:len(want)]
            assert got == want, (cls_name, name, got, want)
            checked += 1
    assert params(pkg.get_gt_and_K) == params(r_cam.get_gt_and_K)
    assert np.array_equal(pkg.DEFAULT_K, r_cam.DEFAULT_K) and pkg.DEFAULT_K.dtype == r_cam.DEFAULT_K.dtype
    assert pkg.SYMMETRIC_OBJECT_IDS == r_add.SYMMETRIC_OBJECT_IDS
    assert checked >= 12


def test_drop_in_import_layout():
    """`sys.path.insert(0, <package dir>)` makes the reference's own import lines resolve here."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); "
            "from models.add_loss import ADDLoss, SYMMETRIC_OBJECT_IDS; from models.pose_loss import PoseLoss; "
            "from utils.camera import DEFAULT_K, get_gt_and_K; import utils; "
            "print(ADDLoss.__module__, PoseLoss.__module__, float(DEFAULT_K[0,0]))"
            % os.path.join(REPO, "6d-pose-estimation_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["models.add_loss", "models.pose_loss", "572.4114"]


def test_default_k_and_gt_reader(pkg, tmp_path):
    import yaml
    g = load_golden("pinhole")
    assert pkg.DEFAULT_K.dtype == np.float64 and np.array_equal(pkg.DEFAULT_K, g["default_K"])
    d = tmp_path / "data" / "05"
    d.mkdir(parents=True)
    K7 = [500.0, 0, 320.0, 0, 501.0, 240.0, 0, 0, 1.0]
    yaml.safe_dump({3: {"cam_K": K7, "depth_scale": 1.0}}, open(d / "info.yml", "w"))
    R = [float(i) for i in range(9)]
    yaml.safe_dump({3: [{"obj_id": 2, "cam_R_m2c": R, "cam_t_m2c": [1, 2, 3]},
                        {"obj_id": 5, "cam_R_m2c": R, "cam_t_m2c": [100.0, 200.0, 900.0], "obj_bb": [1, 2, 3, 4]}]},
                   open(d / "gt.yml", "w"))
    r, t, K = pkg.get_gt_and_K(str(tmp_path / "data"), "05", 3)
    assert np.array_equal(K, np.array(K7).reshape(3, 3)) and np.array_equal(r, np.array(R).reshape(3, 3))
    assert np.allclose(t, [0.1, 0.2, 0.9])
    r, t, K = pkg.get_gt_and_K(str(tmp_path / "data"), "05", 9)        # frame missing: first K, no pose
    assert r is None and t is None and K[0, 0] == 500.0
    r, t, K = pkg.get_gt_and_K(str(tmp_path / "data"), "06", 0)        # nothing on disk: DEFAULT_K copy
    assert r is None and t is None and np.array_equal(K, pkg.DEFAULT_K) and K is not pkg.DEFAULT_K


def test_loader_reproduces_reference(pkg, tmp_path):
    """ADDLoss(model_dir, 'cpu') parses the same PLY bytes to the same points/diameters
    as the reference under the same np.random.seed (face-line quirk, outlier filter,
    diameter fallbacks, 500-point cap).  Loading is host-only, so 'cpu' is allowed here."""
    g = load_golden("loader")
    for name, text in zip(g["file_names"], g["file_texts"]):
        (tmp_path / str(name)).write_text(str(text))
    np.random.seed(1234)
    crit = pkg.ADDLoss(str(tmp_path), "cpu")
    assert sorted(crit.points) == [int(i) for i in g["ids"]]
    for i, d in zip(g["ids"], g["dia"]):
        assert crit.diameters[int(i)] == d
        assert same_bits(crit.points[int(i)].numpy(), g[f"pts_{i}"])
    assert crit.points[0].shape[0] == 500 and crit.points[0].dtype == torch.float32


def test_workloads_are_deterministic(W):
    a, b = W.config2_chunk(3), W.config2_chunk(3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    full = W.config2(8192)
    assert np.array_equal(full[0][4096:], W.config2_chunk(1)[0])
    g = load_golden("eval_cfg1")
    pts, dia, poses = W.config1(seed=1)
    assert np.array_equal(pts[0], g["mesh_0"]) and np.array_equal(poses[0], g["pq"])


def test_dropin_launcher_merges_reference_packages(tmp_path):
    """dropin.install(): hot-path modules from this repo, everything else from a reference
    tree that inserts itself at sys.path[0] (like the reference's scripts do)."""
    import subprocess, sys
    ref = tmp_path / "ref"
    (ref / "models").mkdir(parents=True); (ref / "utils").mkdir(); (ref / "scripts").mkdir()
    (ref / "models" / "__init__.py").write_text("")
    (ref / "models" / "add_loss.py").write_text("class ADDLoss: ORIGIN = 'reference'\n")
    (ref / "models" / "pose_net_rgb.py").write_text("class PoseNetRGB: ORIGIN = 'reference'\n")
    (ref / "models" / "pose_net_rgb_geometric.py").write_text(
        "class PoseNetRGBGeometric:\n    def _compute_pinhole_translation(self, z, c, K): return 'reference'\n")
    (ref / "models" / "pose_net_rgbd_geometric.py").write_text(
        "class PoseNetRGBDGeometric:\n    def _compute_pinhole_translation(self, d, c, K): return 'reference'\n")
    (ref / "utils" / "__init__.py").write_text("")
    (ref / "utils" / "extra_tool.py").write_text("def where(): return 'reference'\n")
    (ref / "scripts" / "s.py").write_text(
        "import os, sys\nPROJECT_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))\n"
        "sys.path.insert(0, PROJECT_ROOT)\n"
        "from models.add_loss import ADDLoss\nfrom models.pose_loss import PoseLoss\n"
        "from models.pose_net_rgb import PoseNetRGB\nfrom models.pose_net_rgb_geometric import PoseNetRGBGeometric\n"
        "from utils.mesh_utils import load_mesh_corners\nfrom utils.visualization import project_points\n"
        "from utils.camera import DEFAULT_K\nfrom utils import draw_3d_box\nfrom utils.extra_tool import where\n"
        "print(ADDLoss.__module__, getattr(ADDLoss, 'ORIGIN', 'b200'), PoseNetRGB.ORIGIN, where(),"
        " PoseNetRGBGeometric._compute_pinhole_translation.__name__, float(DEFAULT_K[1, 1]))\n")
    launcher = os.path.join(REPO, "6d-pose-estimation_b200", "dropin.py")
    out = subprocess.run([sys.executable, launcher, str(ref), "scripts/s.py"], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["models.add_loss", "b200", "reference", "reference", "<lambda>", "573.57043"]


def test_mesh_corners_and_projection_match_reference(pkg, tmp_path):
    """N4 host helpers against the reference's utils/mesh_utils.py and utils/visualization.py
    (golden vectors from oracle/gen_golden.py)."""
    import importlib
    u = importlib.import_module("6d-pose-estimation_b200.utils")
    g = load_golden("projection")
    (tmp_path / "obj_07.ply").write_text(str(g["ply_text"]))
    assert np.array_equal(u.load_mesh_corners(str(tmp_path), "07"), g["corners"])
    assert u.load_mesh_corners(str(tmp_path), "08") is None
    for b in range(len(g["uv"])):
        assert np.array_equal(u.project_points(g["corners"], g["quat"][b], g["trans"][b], g["K"]), g["uv"][b])
        assert np.array_equal(u.project_points(g["corners"], g["Rmat"][b], g["trans"][b], g["K"]), g["uv"][b])
    img = np.zeros((480, 640, 3), np.uint8)
    uv = np.clip(g["uv"][5], 0, 400)
    u.draw_3d_box(img, uv); u.draw_axes(img, g["quat"][5], g["trans"][5] + [0, 0, 1.0], g["K"])
    assert img.any()
