"""2-GPU NCCL test of the data-parallel training step (needs >= 2 devices: `gpurun --gpus 2`; skipped on a
single-GPU box): the all-reduced flat gradient equals the sum of the per-rank gradients, and both ranks hold
bit-identical parameters after the update."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import yolov3_tensorflow_b200 as pkg
    from yolov3_tensorflow_b200 import parallel
    from oracle import yolov3_oracle as O
    from tests.test_gpu_path import _train_case
    parallel.init_from_env("nccl")
    params, x, y_true = _train_case(n=4)
    lo, hi = parallel.shard_batch(4, rank, world)
    xs = torch.from_numpy(x[lo:hi]).cuda()
    ys = [torch.from_numpy(y[lo:hi]).cuda() for y in y_true]
    # local backward (lr 0, no collective), then the data-parallel tail exactly as train_step runs it:
    # ONE all-reduce of the flat gradient buffer, mean factor folded into the fused optimizer kernel.
    # (Two separate backward runs are not comparable bit-for-bit: the BN statistics are accumulated with
    #  fp32 atomics, and at random init this network's BN backward amplifies that last-bit noise.)
    from yolov3_tensorflow_b200._lib import lib, check, stream_handle
    m = pkg.yolov3(80, O.COCO_ANCHORS, dtype="bf16")
    m.set_params(params, "HWIO")
    m.train_step(xs, ys, 0.0, data_parallel=False)
    plan = m._last_plan
    g_local = plan.grad_flat().clone()
    gs = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gs, g_local)
    scale = parallel.allreduce_gradients(plan.grad_flat())
    assert scale == 1.0 / world
    ref = gs[0] + gs[1]
    err = float((plan.grad_flat() - ref).abs().max() / ref.abs().max())
    # momentum 0: the lr-0 local step above left each rank's LOCAL gradient in its velocity buffers
    check(lib.yb_net_train_update(plan.handle, 1e-3, scale, 0.0, 5e-4, 100.0, stream_handle()), "update")
    w = torch.cat([plan.conv_params(i)["w"].reshape(-1) for i in (0, 30, 74)])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same = bool(torch.equal(ws[0], ws[1])) and not bool(torch.equal(w, torch.cat([torch.from_numpy(
        np.ascontiguousarray(np.transpose(params[i]["w"], (3, 0, 1, 2)))).reshape(-1) for i in (0, 30, 74)]).cuda()))
    # and the public API end to end: a DP train_step leaves both ranks with identical parameters
    m2 = pkg.yolov3(80, O.COCO_ANCHORS, dtype="bf16")
    m2.set_params(params, "HWIO")
    m2.train_step(xs, ys, 1e-3)
    w2 = torch.cat([m2._last_plan.conv_params(i)["w"].reshape(-1) for i in (0, 30, 74)])
    ws2 = [torch.empty_like(w2) for _ in range(world)]
    dist.all_gather(ws2, w2)
    same = same and bool(torch.equal(ws2[0], ws2[1]))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, err, same))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_train_step_nccl_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(2))
    for rank, err, same in res:
        assert err < 1e-5, f"rank {rank}: all-reduced gradient differs from the sum of rank gradients ({err})"   # wgrad atomics: order noise only
        assert same, "ranks hold different parameters after the data-parallel update"
