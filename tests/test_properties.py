"""Size-independent properties of the path (hypothesis on the CPU oracle; the GPU twin of
each property lives in test_gpu_parity.py::test_config2_full_size_properties and below)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import same_bits


def _case(W, n, batch, seed, rot_sigma):
    pts = {0: W.sphere_mesh(n, 0.1, seed), 9: W.box_mesh(n, (0.1, 0.12, 0.05), seed + 1)}
    dia = {0: 0.1, 9: 0.1646}
    pq, pt, gq, gt = W.random_poses(batch, seed + 2, rot_sigma=rot_sigma)
    obj = np.where(np.arange(batch) % 2 == 0, 0, 9).astype(np.int64)
    return pts, dia, pq, pt, gq, gt, obj


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 300), seed=st.integers(0, 10_000), rot=st.sampled_from([1e-3, 0.05, 0.4]))
def test_oracle_invariants(oracle, W, n, seed, rot):
    pts, dia, pq, pt, gq, gt, obj = _case(W, n, 6, seed, rot)
    t = oracle.MeshTable(pts, dia)
    add, adds, hit, valid = oracle.add_eval(t, pq, pt, gq, gt, obj)
    assert valid.all() and np.all(adds <= add) and np.all(add >= 0)          # min over j includes j == i
    # q and -q are the same rotation, and _quat_to_mat only uses products of two components
    a2 = oracle.add_eval(t, -pq, pt, -gq, gt, obj)
    assert same_bits(a2[0], add) and same_bits(a2[1], adds) and np.array_equal(a2[2], hit)
    # prediction == ground truth: every distance is exactly zero and the pose is accepted
    z = oracle.add_eval(t, gq, gt, gq, gt, obj)
    assert not z[0].any() and not z[1].any() and z[2].all()
    # the decision is the float64 compare of the returned float32 distance
    thr = np.where(obj == 9, 0.1 * dia[9], 0.1 * dia[0])
    eff = np.where(obj == 9, adds, add).astype(np.float64)
    assert np.array_equal(hit.astype(bool), eff < thr)


@settings(max_examples=25, deadline=None)
@given(batch=st.integers(1, 40), seed=st.integers(0, 10_000), wr=st.floats(0.1, 5), wt=st.floats(0.1, 20))
def test_pose_loss_invariants(oracle, W, batch, seed, wr, wt):
    pq, pt, gq, gt = W.random_poses(batch, seed, rot_sigma=0.3, trans_sigma=0.05)
    o = oracle.pose_loss(pq, pt, gq, gt, wr, wt, "geodesic")
    assert 0.0 <= float(o["rot"]) <= np.pi + 1e-5 and float(o["trans"]) >= 0.0
    assert np.isclose(float(o["loss"]), np.float32(wr) * o["rot"] + np.float32(wt) * o["trans"], rtol=1e-6)
    # double cover: the loss does not see the sign of either quaternion; scale of the raw head output neither
    o2 = oracle.pose_loss(-3.0 * pq, pt, gq, gt, wr, wt, "geodesic")
    assert np.isclose(float(o2["loss"]), float(o["loss"]), rtol=1e-5)
    # the rotation gradient is tangent: a normalised quaternion cannot move along itself
    dots = np.abs((o["grad_q"] * pq).sum(1))
    assert np.all(dots <= 1e-5 * np.maximum(np.abs(o["grad_q"]).max(1), 1e-30) + 1e-12)
    # identical inputs: zero loss and zero rotation gradient
    z = oracle.pose_loss(gq, gt, gq, gt, wr, wt, "geodesic")
    assert float(z["loss"]) == 0.0 and not z["grad_q"].any() and not z["grad_t"].any()
