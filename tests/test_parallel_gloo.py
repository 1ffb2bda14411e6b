"""`not gpu`: the data-parallel host logic (parallel.py) on 2 CPU ranks over gloo: gradient buckets derived from the
lowered backward call list cover the flat gradient arena exactly once, are ordered by when the gradient becomes
final (shared hourglass weights first, stem last), and the bucketed all-reduce leaves the MEAN on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import progressive_process_for_human_pose_estimation_b200.try_with_torch as twt
from progressive_process_for_human_pose_estimation_b200 import _lib as L
from progressive_process_for_human_pose_estimation_b200.parallel import DataParallel, GradientReducer
from progressive_process_for_human_pose_estimation_b200.plan import Builder, Plan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build_plan():
    twt.nStack, twt.nOutChannels = 2, 16
    torch.manual_seed(0)
    net = twt.creatModel()
    b = Builder(True, True)
    net._emit(b, b.input_image(2, 256, 256))
    return net, Plan(b, list(net.named_parameters()), torch.device("cpu"), torch.bfloat16)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if not os.path.exists(L.LIB_PATH):
            L.build()
        net, plan = _build_plan()
        segs = plan.plan_gradient_buckets()
        # (1) segments tile the call list, ranges tile the arena exactly once
        assert segs[0][0] == 0 and segs[-1][1] == len(plan.bwd_calls)
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        cover = torch.zeros(plan.grad_arena.numel(), dtype=torch.int32)
        for _, _, ranges in segs:
            for lo, hi in ranges:
                cover[lo:hi] += 1
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        # (2) stem weights become final last; the shared hourglass weights before them
        names = [n for n, _ in plan.params]
        stem_off = plan.grad_offsets[names.index("conv1.weight")]
        last_ranges = segs[-1][2]
        assert any(lo <= stem_off < hi for lo, hi in last_ranges)
        hg_off = plan.grad_offsets[names.index("hourglass1.residual_block.conv2.weight")]
        assert len(segs) >= 2 and not any(lo <= hg_off < hi for lo, hi in last_ranges)
        # (3) bucketed all-reduce == mean over ranks, bucket by bucket
        g = torch.Generator().manual_seed(100 + rank)
        plan.grad_arena.copy_(torch.randn(plan.grad_arena.numel(), generator=g))
        mine = plan.grad_arena.clone()
        red = GradientReducer()
        for _, _, ranges in segs:
            red.reduce_async(plan.grad_arena, ranges)
        red.wait()
        others = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(others, mine)
        expect = sum(others) / world
        assert torch.allclose(plan.grad_arena, expect, rtol=0, atol=1e-6)
        assert red.calls == sum(len(r) for _, _, r in segs)
        # (4) DataParallel broadcasts rank 0's parameters and buffers
        torch.manual_seed(rank)
        other = twt.creatModel()
        dp = DataParallel(other)
        ref = [torch.empty_like(other.conv2.weight) for _ in range(world)]
        dist.all_gather(ref, other.conv2.weight.data)
        assert torch.equal(ref[0], ref[1])
        assert list(dp.state_dict().keys()) == list(other.state_dict().keys())
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()
        twt.nStack, twt.nOutChannels = 4, 17


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
