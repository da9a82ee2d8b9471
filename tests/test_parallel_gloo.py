"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (the NCCL path runs the same
code on the GPU box): batch sharding, the gradient all-reduce + mean factor, max-over-ranks timing, and the
data-parallel equivalence rule of SURVEY.md §8e checked on the CPU oracle (2 ranks x n images == 1 rank x 2n)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from yolov3_tensorflow_b200 import parallel
    from oracle import yolov3_oracle as O
    from tests.synth import gen_fms
    r, w, l = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    lo, hi = parallel.shard_batch(4, rank, world)
    assert (lo, hi) == (2 * rank, 2 * rank + 2)
    with pytest.raises(ValueError):
        parallel.shard_batch(5, rank, world)
    # ---- gradient all-reduce: DP equivalence on the loss gradient of the oracle ----
    n, h, wd, cn = 4, 64, 64, 4
    f = gen_fms(7, n, h, wd, cn, scale=1.0)
    rng = np.random.default_rng(5)
    ys = [[], [], []]
    for i in range(n):
        boxes, labels = O.synth_gt(rng, wd, h, cn, 6)
        y = O.process_box(boxes, labels, [wd, h], cn, O.COCO_ANCHORS)
        for j in range(3):
            ys[j].append(y[j])
    y_true = [np.stack(y) for y in ys]
    # single-device reference: the whole batch
    _, g_full = O.loss_and_grad(f, y_true, O.COCO_ANCHORS, (h, wd), cn, dtype=torch.float64)
    # this rank: its shard only (each loss term is a mean over the LOCAL batch)
    _, g_loc = O.loss_and_grad([a[lo:hi] for a in f], [a[lo:hi] for a in y_true], O.COCO_ANCHORS, (h, wd), cn, dtype=torch.float64)
    # "flat gradient buffer" of a rank = d(local loss)/d(shared parameter); emulate a shared parameter theta with
    # fm = theta * f : d/dtheta = sum(g * f)
    flat = torch.tensor([float((g * a[lo:hi]).sum()) for g, a in zip(g_loc, f)], dtype=torch.float64)
    scale = parallel.allreduce_gradients(flat)
    assert scale == 0.5
    ref = torch.tensor([float((g * a).sum()) for g, a in zip(g_full, f)], dtype=torch.float64)
    assert torch.allclose(flat * scale, ref, rtol=1e-9, atol=1e-12), (flat * scale, ref)
    # ---- bucketed reduction: buckets listed heads-first, every element reduced exactly once ----
    sizes = [5, 0, 7, 3, 9, 1]
    buckets = parallel.gradient_buckets(sizes, 8)
    assert buckets == [(4, 5), (2, 3), (0, 1)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    flat2 = torch.arange(float(sum(sizes)), dtype=torch.float64) * (rank + 1)
    red = parallel.BucketedAllReduce()
    for lo_l, hi_l in buckets:
        red.reduce(flat2[offs[lo_l]: offs[hi_l + 1]])
    assert red.wait() == 0.5
    assert torch.equal(flat2, torch.arange(float(sum(sizes)), dtype=torch.float64) * 3.0)
    # ---- timing reduction ----
    assert parallel.max_over_ranks(10.0 + rank, "cpu") == 11.0
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, True))


def test_data_parallel_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]


def test_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    from yolov3_tensorflow_b200 import parallel
    t = torch.ones(4)
    assert parallel.allreduce_gradients(t) == 1.0 and torch.equal(t, torch.ones(4))
    assert parallel.max_over_ranks(3.0, "cpu") == 3.0
    assert parallel.shard_batch(8, 0, 1) == (0, 8)
    assert parallel.gradient_buckets([4, 4, 4], 100) == [(0, 2)] and parallel.gradient_buckets([4, 4, 4], 1) == [(2, 2), (1, 1), (0, 0)]
    red = parallel.BucketedAllReduce()
    red.reduce(t)
    assert red.wait() == 1.0 and torch.equal(t, torch.ones(4))
