"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path (hypothesis
sharding + the per-object count all-reduce).  No kernels run here; per-pose results come
from the golden vectors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import importlib, sys
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("6d-pose-estimation_b200")
    g = np.load(os.path.join(REPO, "tests", "golden", "eval_cfg1_mixed.npz"))
    B = len(g["obj"])
    lo, hi = pkg.shard_range(B, rank, world)
    acc = pkg.sweep.Accumulators(1, 16, "cpu")
    for i in range(lo, hi):                       # what the kernel's atomics do on a GPU
        if g["valid"][i]:
            o = int(g["obj"][i])
            acc.valid[0, o] += 1
            acc.hits[0, o] += int(g["hit"][i])
            acc.add_sum[0, o] += float(g["add"][i])
            acc.adds_sum[0, o] += float(g["adds"][i])
    acc.all_reduce()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), hits=acc.hits.numpy(), valid=acc.valid.numpy(),
             add=acc.add_sum.numpy(), adds=acc.adds_sum.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly(pkg):
    for total in (0, 1, 7, 64, 65536, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [pkg.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.shard_range(10, 2, 2)


def test_count_allreduce_world2_gloo(pkg, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = load_golden("eval_cfg1_mixed")
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ("hits", "valid", "add", "adds"):
        assert np.array_equal(r0[k], r1[k])                     # every rank holds the reduced result
    v = g["valid"].astype(bool)
    for o in np.unique(g["obj"][v]):
        sel = v & (g["obj"] == o)
        assert r0["valid"][0, o] == sel.sum() and r0["hits"][0, o] == g["hit"][sel].sum()
        assert np.isclose(r0["add"][0, o], g["add"][sel].astype(np.float64).sum(), rtol=1e-13)
    # accuracy from the reduced integers == the reference's aggregate over the whole batch
    acc = 100.0 * r0["hits"].sum() / r0["valid"].sum()
    assert acc == g["agg"][2]


def test_reference_batch_means(pkg):
    g = load_golden("eval_cfg1_mixed")
    # one batch == eval_metrics of the whole batch (golden 'agg' came from the reference)
    m = pkg.reference_batch_means(g["add"], g["adds"], g["hit"], g["valid"], batch_size=len(g["obj"]))
    assert np.array_equal([m["add_mean"], m["add_s_mean"], m["add_01d_acc"]], g["agg"])
    # batches of 16 as compare_all_models.py:121 -> mean of the two per-batch dicts
    m16 = pkg.reference_batch_means(g["add"], g["adds"], g["hit"], g["valid"], batch_size=16)
    halves = [pkg.reference_batch_means(g["add"][s], g["adds"][s], g["hit"][s], g["valid"][s], 16)
              for s in (slice(0, 16), slice(16, 32))]
    assert np.isclose(m16["add_mean"], (halves[0]["add_mean"] + halves[1]["add_mean"]) / 2, rtol=1e-15)
    assert m16["add_01d_acc"] == (halves[0]["add_01d_acc"] + halves[1]["add_01d_acc"]) / 2
