"""Batched, device-resident evaluation sweeps (SURVEY.md section 8e and "next" row N2).

The reference evaluates pose hypotheses in a Python loop inside
``ADDLoss.eval_metrics`` and aggregates on the host (``scripts/visualization/
compare_all_models.py:65-104``: batches of 16, mean of per-batch means).  This module is
the caller-side counterpart of the kernels:

* ``PoseEvaluator`` keeps per-object accumulators (ADD-0.1d hits, valid poses, float64
  ADD / ADD-S sums) on the device; every ``evaluate`` call is one kernel launch and no
  host synchronisation;
* hypotheses are independent, so N GPUs shard the hypothesis axis (``shard_range``) with
  no exchange during compute; the only collective is one all-reduce of the integer
  accumulators (plus the float64 sums) at the end (``PoseEvaluator.all_reduce``) --
  accuracy = hits / valid is computed from integers and is therefore identical for any
  number of GPUs;
* ``reference_batch_means`` reproduces the reference's reporting convention (mean of
  per-batch means) from per-pose outputs, for scripts that must print the same numbers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as core
from .models.add_loss import SYMMETRIC_OBJECT_IDS

VARIANTS = ("rgb", "rgb_geometric", "rgbd", "rgbd_geometric")


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `total` items for `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reference_batch_means(add, add_s, hit, valid, batch_size: int) -> dict:
    """The numbers ``compare_all_models.py`` prints (:95-104): per batch the reference
    takes ``eval_metrics`` (float64 means over the valid poses of the batch, 0 for an
    empty batch) and then averages the per-batch values over the number of batches --
    the last, shorter batch is over-weighted.  Inputs are per-pose host arrays."""
    add, add_s = np.asarray(add, np.float32), np.asarray(add_s, np.float32)
    hit, valid = np.asarray(hit).astype(bool), np.asarray(valid).astype(bool)
    n = add.shape[0]
    sums = np.zeros(3, np.float64)
    batches = 0
    for lo in range(0, n, batch_size):
        v = valid[lo:lo + batch_size]
        batches += 1
        if not v.any():
            continue
        sums[0] += np.mean(add[lo:lo + batch_size][v].astype(np.float64)) * 1000
        sums[1] += np.mean(add_s[lo:lo + batch_size][v].astype(np.float64)) * 1000
        sums[2] += np.mean(hit[lo:lo + batch_size][v].astype(np.float64)) * 100
    if batches == 0:
        return {"add_mean": 0.0, "add_s_mean": 0.0, "add_01d_acc": 0.0}
    return {"add_mean": sums[0] / batches, "add_s_mean": sums[1] / batches, "add_01d_acc": sums[2] / batches}


class Accumulators:
    """[n_rows, n_slots] device accumulators; one row per model variant (or just one)."""

    def __init__(self, n_rows: int, n_slots: int, device):
        self.hits = torch.zeros(n_rows, n_slots, dtype=torch.int64, device=device)
        self.valid = torch.zeros(n_rows, n_slots, dtype=torch.int64, device=device)
        self.add_sum = torch.zeros(n_rows, n_slots, dtype=torch.float64, device=device)
        self.adds_sum = torch.zeros(n_rows, n_slots, dtype=torch.float64, device=device)

    def row(self, r: int):
        return [self.hits[r], self.valid[r], self.add_sum[r], self.adds_sum[r]]

    def zero_(self):
        for t in (self.hits, self.valid, self.add_sum, self.adds_sum):
            t.zero_()

    def all_reduce(self, group=None):
        """Sum over ranks.  Integers first (exact, order-independent), then the float64
        sums.  Works with NCCL (CUDA tensors) and, for the CPU tests, gloo."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self
        counts = torch.stack([self.hits, self.valid])
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        self.hits.copy_(counts[0]); self.valid.copy_(counts[1])
        sums = torch.stack([self.add_sum, self.adds_sum])
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        self.add_sum.copy_(sums[0]); self.adds_sum.copy_(sums[1])
        return self

    def table(self, row_names=None) -> dict:
        """Host summary: per row and object id -> accuracy %, mean ADD / ADD-S in mm, counts."""
        hits, valid = self.hits.cpu().numpy(), self.valid.cpu().numpy()
        a, s = self.add_sum.cpu().numpy(), self.adds_sum.cpu().numpy()
        out = {}
        for r in range(hits.shape[0]):
            name = row_names[r] if row_names else r
            rows = {}
            for o in np.nonzero(valid[r])[0]:
                n = int(valid[r, o])
                rows[int(o)] = {"n": n, "hits": int(hits[r, o]), "add_01d_acc": 100.0 * int(hits[r, o]) / n,
                                "add_mean": a[r, o] / n * 1000, "add_s_mean": s[r, o] / n * 1000}
            tot = int(valid[r].sum())
            rows["all"] = {"n": tot, "hits": int(hits[r].sum()),
                           "add_01d_acc": 100.0 * int(hits[r].sum()) / tot if tot else 0.0}
            out[name] = rows
        return out


class PoseEvaluator:
    """Device-resident ADD / ADD-S / ADD-0.1d evaluation over large hypothesis sets."""

    def __init__(self, points: dict, diameters: dict, device, symmetric_ids=SYMMETRIC_OBJECT_IDS,
                 n_rows: int = 1):
        self.device = core.require_cuda(device)
        self.table = core.MeshTable(points, diameters, symmetric_ids, self.device)
        self.acc = Accumulators(n_rows, self.table.n_slots, self.device)
        self.launches = 0

    def evaluate(self, pred_q, pred_t, gt_q, gt_t, obj_ids, row: int = 0, want_adds: bool = True,
                 sort: bool = True, per_pose: bool = False):
        """One launch over B poses; accumulates into row `row`.  No host sync unless
        per_pose outputs are converted by the caller."""
        dev = self.device
        pq = core.as_cuda_f32(pred_q, dev, (4,)); pt = core.as_cuda_f32(pred_t, dev, (3,))
        gq = core.as_cuda_f32(gt_q, dev, (4,)); gt = core.as_cuda_f32(gt_t, dev, (3,))
        obj = obj_ids.to(dev, torch.int64).reshape(-1).contiguous()
        order = torch.argsort(obj, stable=True).to(torch.int32) if sort else None
        add, adds, hit, valid, _ = self.table.evaluate(pq, pt, gq, gt, obj, want_adds, order, self.acc.row(row))
        self.launches += 1
        return (add, adds, hit, valid) if per_pose else None

    def all_reduce(self, group=None):
        return self.acc.all_reduce(group)


def synth_chunk(n: int, seed: int, device, rot_sigma=0.05, trans_sigma=0.005):
    """Seeded hypotheses generated ON the device (config 5 is 52 M poses; host generation
    would dominate).  Same distributions as workloads.random_poses."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    gq = torch.nn.functional.normalize(torch.randn(n, 4, generator=g, device=device), dim=1)
    pq = torch.nn.functional.normalize(gq + rot_sigma * torch.randn(n, 4, generator=g, device=device), dim=1)
    u = torch.rand(n, 3, generator=g, device=device)
    gt = torch.stack([u[:, 0] * 0.4 - 0.2, u[:, 1] * 0.4 - 0.2, u[:, 2] * 0.8 + 0.4], 1)
    pt = gt + trans_sigma * torch.randn(n, 3, generator=g, device=device)
    return pq, pt, gq, gt


def variant_translation(variant: str, pred_t, gt_t, K, seed: int):
    """How each model family produces its translation:
    rgb / rgbd regress xyz directly; rgb_geometric predicts Z and back-projects the bbox
    centre (kernel d1); rgbd_geometric reads Z from a depth crop under the crop-space
    centre (kernel d2)."""
    from .utils.camera import depth_backproject, pinhole_translation
    if variant in ("rgb", "rgbd"):
        return pred_t
    dev = pred_t.device
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    uv = torch.stack([gt_t[:, 0] / gt_t[:, 2] * fx + cx, gt_t[:, 1] / gt_t[:, 2] * fy + cy], 1)
    if variant == "rgb_geometric":
        return pinhole_translation(pred_t[:, 2:3].contiguous(), uv, K)
    if variant == "rgbd_geometric":
        # a tiny synthetic "crop": 8x8 depth patch around the centre, noisy sensor depth of the GT
        n = pred_t.shape[0]
        g = torch.Generator(device=dev); g.manual_seed(int(seed) + 7)
        depth = gt_t[:, 2].reshape(n, 1, 1) + 0.002 * torch.randn(n, 8, 8, generator=g, device=dev)
        centre = torch.full((n, 2), 3.5, device=dev)
        Kc = K.clone().reshape(1, 3, 3).repeat(n, 1, 1)
        Kc[:, 0, 2] = 3.5 - (uv[:, 0] - cx)          # crop-space principal point so that (u-cx) is preserved
        Kc[:, 1, 2] = 3.5 - (uv[:, 1] - cy)
        return depth_backproject(depth.contiguous(), centre, Kc, clamp_hi=7.0)
    raise ValueError(f"unknown variant {variant!r}")


def evaluate_sweep(points: dict, diameters: dict, device, n_per_block: int, variants=VARIANTS, chunk: int = 65536,
                   seed: int = 5000, rank: int = 0, world: int = 1, group=None, K=None):
    """compare_all_models-style sweep: for every (object, variant) block, `n_per_block`
    seeded hypotheses, the hypothesis axis of each block sliced across `world` ranks.
    Returns (Accumulators after the all-reduce, number of kernel launches on this rank)."""
    dev = core.require_cuda(device)
    ev = PoseEvaluator(points, diameters, dev, n_rows=len(variants))
    if K is None:
        from .utils.camera import DEFAULT_K
        K = torch.tensor(DEFAULT_K, dtype=torch.float32, device=dev)
    objs = sorted(points)
    lo, hi = shard_range(n_per_block, rank, world)
    for oi, oid in enumerate(objs):
        for vi, var in enumerate(variants):
            # chunks live on a global grid and are seeded by their global index, so the data
            # (and therefore every count) is the same for any number of ranks
            for c in range(lo // chunk, (hi + chunk - 1) // chunk):
                c0 = c * chunk
                n = min(chunk, n_per_block - c0)
                cseed = seed + 1_000_003 * oi + 10_007 * vi + c
                pq, pt, gq, gt = synth_chunk(n, cseed, dev)
                pt = variant_translation(var, pt, gt, K, cseed)
                a, b = max(lo, c0) - c0, min(hi, c0 + n) - c0
                if b <= a:
                    continue
                obj = torch.full((b - a,), oid, dtype=torch.int64, device=dev)
                ev.evaluate(pq[a:b], pt[a:b], gq[a:b], gt[a:b], obj, row=vi, sort=False)
    ev.all_reduce(group)
    return ev.acc, ev.launches
