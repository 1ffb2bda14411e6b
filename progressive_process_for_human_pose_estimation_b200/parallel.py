"""Data parallelism for the hourglass hot path: one process per GPU, batch sharded per rank, BatchNorm statistics
per rank (the reference has no SyncBN), one exchange per step: all-reduce(mean) of the parameter gradients over
NCCL / NVLink, bucketed by *when the gradient becomes final* and overlapped with the rest of the backward pass.

The reference is single-GPU; this is the B200-native addition north_star asks for.  Because every
hourglass weight is shared by all stacks (try_with_torch.py:268-273,285-297), its gradient is only complete when
the first stack's backward has run; only the stem's backward is left to hide the transfer (7.6 MB fp32 total).
The plan therefore splits its backward call list at that point (Plan.plan_gradient_buckets): bucket 0 (hourglass
+ head, 98 % of the bytes) is all-reduced on a side stream while the stem's backward runs, bucket 1 (stem)
follows.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
import torch
import torch.distributed as dist


class GradientReducer:
    """Averages ranges of a flat fp32 gradient buffer across ranks on a side stream."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = None
        self.calls = 0

    def reduce_async(self, flat, ranges):
        if self.world == 1:
            return
        ranges = ranges if ranges else [(0, flat.numel())]
        if flat.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=flat.device)
            self.stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self.stream):
                for lo, hi in ranges:
                    part = flat[lo:hi]
                    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
                    part.mul_(1.0 / self.world)
                    self.calls += 1
        else:
            for lo, hi in ranges:
                part = flat[lo:hi]
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
                part.mul_(1.0 / self.world)
                self.calls += 1

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


class DataParallel(torch.nn.Module):
    """Wrap a drop-in model for data-parallel training:

        dist.init_process_group("nccl"); torch.cuda.set_device(local_rank)
        model = hg.parallel.DataParallel(m.creatModel().cuda())
        loss = sum(mse(r, y) for r in model(x_shard)); loss.backward(); opt.step()

    Parameters (and BatchNorm buffers) are broadcast from rank 0 at construction; after backward() every rank
    holds the mean gradient.  `state_dict()` keys are those of the wrapped module.
    """

    def __init__(self, module, group=None, broadcast=True):
        super().__init__()
        self.module = module
        self.reducer = GradientReducer(group)
        if broadcast and self.reducer.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, 0, group=group)

    def forward(self, x):
        plan = self.module._plan_for(x)
        if plan.reducer is None and self.reducer.world > 1 and plan.need_bwd:
            plan.reducer = self.reducer
            plan.plan_gradient_buckets()
        return self.module(x)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self.module.load_state_dict(*a, **k)
