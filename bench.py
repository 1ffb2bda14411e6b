#!/usr/bin/env python
"""bench.py -- training throughput of the 8-stack hourglass (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU implementation (oracle port) on the host cores

A step = forward of the weight-shared 8-stack network (try_with_torch.creatModel, nStack=8, 16 heatmaps) on a
batch of 32 images per GPU at 256x256, eight nn.MSELoss terms on Gaussian targets, backward, Adam update.
`value` times that with the batch resident in HBM; `e2e` adds, inside the timed region of every step, the copy of
the step's images+targets from pinned host memory and a device->host read of the loss.  One JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train images/sec, 8-stack hourglass 256x256"
UNIT = "images/s"
NSTACK, NJOINT, IMG = 8, 16, 256
# conv FLOPs (2*MAC) per image of this network, hook-counted on the reference
FWD_GFLOP_PER_IMG = 97.272
TRAIN_GFLOP_PER_IMG = 291.51


def dist_info():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:  # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(B, seed, device):
    """Seeded synthetic batch: N(0,1) images and Gaussian targets rendered by the library's own render kernel
    from random MPII-style annotations (one person, 16 joints, 640x480 image)."""
    import numpy as np

    import progressive_process_for_human_pose_estimation_b200 as hg

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, IMG, IMG, generator=g)
    r = np.random.RandomState(seed + 1)
    kp = np.zeros([B, 1, NJOINT, 3])
    kp[..., 0] = r.randint(0, 640, [B, 1, NJOINT])
    kp[..., 1] = r.randint(0, 480, [B, 1, NJOINT])
    kp[..., 2] = r.randint(0, 3, [B, 1, NJOINT])
    wh = np.tile(np.array([[640.0, 480.0]]), (B, 1))
    y = hg.gaussian_heatmaps(kp, wh, truncate=True, device=device)
    return x, y


_KERNEL_NAMES = {
    "hg_conv_fprop_ex": "conv_gemm_kernel<kPlain> fprop", "hg_conv_fprop_bn": "conv_gemm_kernel<kFold> fprop (BN+ReLU folded)",
    "hg_conv_dgrad": "conv_gemm_kernel<kPlain> dgrad", "hg_conv_dgrad_bn": "conv_gemm_kernel<kMask> dgrad (+ReLU mask, BN sums)",
    "hg_conv_wgrad": "conv_wgrad_kernel", "hg_conv_wgrad_bn": "conv_wgrad_kernel<FOLD>",
    "hg_bn_apply": "bn_apply_kernel", "hg_bn_bwd_apply": "bn_bwd_apply_kernel",
}


def kernel_roofline(name, tag, n, t_ms, B, peaks, src, total_ms, traffic_db):
    """Roofline row of one (entry point, shape) of the per-kernel table, or None when it has no simple model.
    Convolutions: 2*M*Cin*Cout*k*k FLOPs (tag 'Cin->Cout kK @HxW'); BatchNorm passes: bf16 tensors read + written."""
    import re

    avg_s = t_ms / n * 1e-3
    row = None
    mconv = re.match(r"(\d+)->(\d+) k(\d+) @(\d+)x(\d+)", tag or "")
    if name.startswith("hg_conv_") and mconv and name in _KERNEL_NAMES:
        ci, co, k, h, w = (int(v) for v in mconv.groups())
        flops = 2.0 * B * h * w * ci * co * k * k
        kname = _KERNEL_NAMES[name]
        # large-map 1x1 fprop / dgrad launches run the persistent kernel (csrc/conv_persist.cu: >= 512 units of 128 pixels),
        # and so do the 3x3 convolutions with 128 output channels (its transposed-accumulator variant)
        if B * h * w >= 512 * 128 and name in ("hg_conv_fprop_ex", "hg_conv_dgrad", "hg_conv_dgrad_bn"):
            gemm_n = (co if name == "hg_conv_fprop_ex" else ci)
            if k == 1:
                kname = kname.replace("conv_gemm_kernel", "conv_persist_kernel")
            elif k == 3 and (gemm_n + 63) // 64 * 64 == 128:
                kname = kname.replace("conv_gemm_kernel", "conv_persist_kernel<TR>")
        if k == 1:
            # 1x1 convolutions are HBM-bound (8.6 GFLOP over >= 100 MB at 64x64): bf16 tensors read + written per launch
            cip, cop = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
            if name == "hg_conv_wgrad":
                chans = cip + cop                                  # x, dy
            elif name.startswith("hg_conv_dgrad"):
                chans = cop + cip + (cip if name == "hg_conv_dgrad_bn" or "+add" in tag else 0)   # dy, dx (+ raw BN input)
            else:
                chans = cip + cop + (cop if "+res" in tag else 0)  # x, y (+ residual)
            nbytes = B * h * w * chans * 2
            gbs = nbytes / avg_s / 1e9
            peak = float(peaks["hbm_gbs"])
            row = {"kernel": f"{kname} {tag} (tcgen05 + TMA)", "bound": "hbm", "achieved": round(gbs, 1), "peak": peak,
                   "unit": "GB/s", "frac": round(gbs / peak, 4), "algorithmic_bytes_per_launch": nbytes,
                   "algorithmic_flops_per_launch": flops, "tflops": round(flops / avg_s / 1e12, 1),
                   "peak_source": f"{src} hbm_gbs"}
        else:
            ach = flops / avg_s / 1e12
            peak = float(peaks["bf16_tflops"])
            row = {"kernel": f"{kname} {tag} (tcgen05 + TMA)", "bound": "tensor", "achieved": round(ach, 2),
                   "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "algorithmic_flops_per_launch": flops,
                   "peak_source": f"{src} bf16_tflops (burst: per-launch timing)"}
    mbn = re.match(r"C(\d+) M(\d+)( \+addend)?", tag or "")
    if name in ("hg_bn_apply", "hg_bn_bwd_apply") and mbn:
        c, m = int(mbn.group(1)), int(mbn.group(2))
        tensors = 2 if name == "hg_bn_apply" else (4 if mbn.group(3) else 3)   # bf16 tensors read + written
        nbytes = tensors * m * ((c + 63) // 64 * 64) * 2
        gbs = nbytes / avg_s / 1e9
        peak = float(peaks["hbm_gbs"])
        row = {"kernel": f"{_KERNEL_NAMES[name]} {tag}", "bound": "hbm", "achieved": round(gbs, 1), "peak": peak,
               "unit": "GB/s", "frac": round(gbs / peak, 4), "algorithmic_bytes_per_launch": nbytes,
               "peak_source": f"{src} hbm_gbs"}
    if row is None:
        return None
    row.update({"traffic": (traffic_db.get(f"{name} {tag}") if B == 32 else None), "launches_per_step": n,
                "avg_us": round(avg_s * 1e6, 2), "share_of_step_kernel_time": round(t_ms / total_ms, 4)})
    return row


def run_ours(args):
    import torch.distributed as dist

    import progressive_process_for_human_pose_estimation_b200 as hg
    import progressive_process_for_human_pose_estimation_b200.try_with_torch as m
    from progressive_process_for_human_pose_estimation_b200 import _lib as L
    from progressive_process_for_human_pose_estimation_b200.parallel import DataParallel

    world, rank, local = dist_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    for opt in os.environ.get("HG_OPTIONS", "").split(","):   # e.g. HG_OPTIONS=single_wave_deep=0,wgrad_smem_kb=110
        if "=" in opt:
            k_, v_ = opt.split("=")
            L.call("hg_set_option", k_.encode(), int(v_))
    if not L.load().hg_device_ok():
        raise SystemExit("bench.py: device is not compute capability 10.x")
    B = args.batch
    warmup = max(3, args.warmup)
    hg.set_compute_dtype(torch.bfloat16)
    m.nStack, m.nOutChannels = NSTACK, NJOINT
    torch.manual_seed(0)
    net = m.creatModel().to(dev)
    model = DataParallel(net) if world > 1 else net
    # loss and optimizer of the training loop (try_with_torch.py:305-308,317,333-344): the library's own kernels
    # (hg.mse_losses: all 8 MSE terms and their gradients in one pass; hg.Adam: one launch for the 199 parameter
    # tensors).  HG_BENCH_STOCK=1 runs the stock nn.MSELoss modules and torch.optim.Adam on the same model instead.
    stock = os.environ.get("HG_BENCH_STOCK", "0") == "1"
    opt = (torch.optim.Adam if stock else hg.Adam)(net.parameters(), lr=1e-4)
    mse = [torch.nn.MSELoss() for _ in range(NSTACK)]
    x_cpu, y = synth_batch(B, 100 + rank, dev)
    x = x_cpu.to(dev)

    def total_loss(out, yb):
        if not stock:
            return hg.mse_losses(out, yb).sum()
        loss = mse[0](out[0], yb)
        for k in range(1, NSTACK):
            loss = loss + mse[k](out[k], yb)
        return loss

    def step(xb, yb):
        out = model(xb)
        loss = total_loss(out, yb)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()   # still inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    l0 = L.launch_count()
    for i in range(warmup):
        loss = step(x, y)
        if i == 0:
            torch.cuda.synchronize()
            launches_per_step = L.launch_count() - l0  # step 0 runs every C-ABI call eagerly
    torch.cuda.synchronize()
    loss0 = loss.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: step(x, y), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- where the step goes: forward / loss+backward / optimizer, CUDA events, same stream -------------
    phases = None
    if True:  # every rank: the backward pass contains the gradient all-reduce
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(3)]
        for e in ev:
            e[0].record()
            out = model(x)
            e[1].record()
            loss = total_loss(out, y)
            opt.zero_grad()
            loss.backward()
            e[2].record()
            opt.step()
            e[3].record()
        torch.cuda.synchronize()
        med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
        phases = {"forward_ms": round(med([e[0].elapsed_time(e[1]) for e in ev]), 3),
                  "loss_backward_ms": round(med([e[1].elapsed_time(e[2]) for e in ev]), 3),
                  "optimizer_ms": round(med([e[2].elapsed_time(e[3]) for e in ev]), 3)}
        if world > 1:
            # the same backward pass WITHOUT the gradient all-reduce (one graph, gradients stay local: timing only):
            # the difference is what the bucketed, overlapped all-reduce still costs the step
            plan_ = list(net.__dict__["_plans"].values())[-1]
            saved_red, plan_.reducer = plan_.reducer, None
            ev2 = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(4)]
            for e in ev2:
                out = model(x)
                loss = total_loss(out, y)
                opt.zero_grad()
                e[0].record()
                loss.backward()
                e[1].record()
            torch.cuda.synchronize()
            plan_.reducer = saved_red
            nocomm = med([e[0].elapsed_time(e[1]) for e in ev2[1:]])
            ev3 = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3)]
            for e in ev3:
                out = model(x)
                loss = total_loss(out, y)
                opt.zero_grad()
                e[0].record()
                loss.backward()
                e[1].record()
            torch.cuda.synchronize()
            withcomm = med([e[0].elapsed_time(e[1]) for e in ev3])
            phases["backward_ms"] = round(withcomm, 3)
            phases["backward_ms_no_allreduce"] = round(nocomm, 3)
            phases["allreduce_exposed_ms"] = round(withcomm - nocomm, 3)

    # ---- end to end: pinned host batch -> device every step, loss read back every step ---------------
    xh = x_cpu.pin_memory()
    yh = y.cpu().pin_memory()
    bufs = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    state = {"i": 0, "last": 0.0}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            bufs[slot][0].copy_(xh, non_blocking=True)
            bufs[slot][1].copy_(yh, non_blocking=True)

    # every step's loss is copied device -> host (pinned, asynchronously) and read by the host one step later, so the
    # host keeps launching ahead like the reference loop (which prints the loss every 50 iterations); the last loss is
    # read inside the timed region by `drain`
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def read_loss(i):
        loss_ready[i % 2].synchronize()
        state["last"] = float(loss_host[i % 2][0])

    def e2e_step():
        i = state["i"]
        if i == 0:
            upload(0)
        torch.cuda.current_stream().wait_stream(copy_stream)
        xb, yb = bufs[i % 2]
        copy_stream.wait_stream(torch.cuda.current_stream())  # the other slot is free once the previous step is queued
        upload((i + 1) % 2)  # next step's batch travels while this step computes
        loss = step(xb, yb)
        if i > 0:
            read_loss(i - 1)                                   # host read of the previous step's loss
        loss_host[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)   # device -> host copy of this step's loss
        loss_ready[i % 2].record()
        state["i"] = i + 1

    def drain():
        read_loss(state["i"] - 1)

    e2e_step()
    drain()
    state["i"] = 0
    ms_e2e = timed(e2e_step, args.steps, finish=drain)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = xh.numel() * 4 + yh.numel() * 4
    # (the prefetch uploads one extra batch at the very end; bytes are counted per step as copied)

    # ---- the same end-to-end step fed with uint8 pixels (next row N1): the host ships [B,256,256,3] bytes, the
    # GPU applies ToTensor + Normalize(0.5, 0.5) (hg.to_tensor_normalize, bit-exact with torchvision) ---------------
    e2e_u8 = None
    if not stock:
        xu8 = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8).pin_memory()
        ubufs = [torch.empty(B, 256, 256, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        ustate = {"i": 0}

        def upload_u8(slot):
            with torch.cuda.stream(copy_stream):
                ubufs[slot].copy_(xu8, non_blocking=True)
                bufs[slot][1].copy_(yh, non_blocking=True)

        def e2e_u8_step():
            i = ustate["i"]
            if i == 0:
                upload_u8(0)
            torch.cuda.current_stream().wait_stream(copy_stream)
            copy_stream.wait_stream(torch.cuda.current_stream())
            upload_u8((i + 1) % 2)
            loss = step(hg.to_tensor_normalize(ubufs[i % 2]), bufs[i % 2][1])
            if i > 0:
                read_loss(i - 1)
            loss_host[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ready[i % 2].record()
            ustate["i"] = i + 1

        def drain_u8():
            read_loss(ustate["i"] - 1)

        e2e_u8_step()
        drain_u8()
        ustate["i"] = 0
        ms_u8 = timed(e2e_u8_step, args.steps, finish=drain_u8)
        e2e_u8 = {"value": round(world * B * args.steps / (ms_u8 / 1e3), 2), "unit": UNIT,
                  "h2d_bytes_per_step": xu8.numel() + yh.numel() * 4, "d2h_bytes_per_step": 4,
                  "ms_per_step": round(ms_u8 / args.steps, 3),
                  "input": "uint8 NHWC pixels; ToTensor + Normalize(0.5, 0.5) on the GPU"}

    # ---- inference leg (BASELINE configs[4]): 4-stack / 17-joint network in eval(), heatmaps of the last stack
    # decoded and scored by the PCKh threshold sweep, all on the device; images/s over all ranks ----------------
    inference = None
    if not args.no_inference:
        import numpy as np

        m.nStack, m.nOutChannels = 4, 17
        torch.manual_seed(1)
        inet = m.creatModel().to(dev).eval()
        r = np.random.RandomState(7 + rank)
        label = torch.zeros(B, 64, 64, dtype=torch.int64)
        for b_ in range(B):
            pos = r.choice(64 * 64, 17, replace=False)
            for j, pp in enumerate(pos):
                label[b_, pp // 64, pp % 64] = j + 1
        label = label.to(dev)
        x0 = r.uniform(5, 40, [B, 2]).astype("float32")
        rect = torch.from_numpy(np.concatenate([x0, x0 + r.uniform(5, 20, [B, 2]).astype("float32")], 1)).to(dev)

        def infer():
            with torch.no_grad():
                out = inet(x)
                return hg.pckh_sweep_counts(out[-1], label, rect, 0)

        for _ in range(3):
            res = infer()
        ms_inf = timed(infer, args.steps)
        inference = {"metric": "inference images/sec, 4-stack hourglass 256x256 (17 joints) + argmax decode + PCKh sweep",
                     "value": round(world * B * args.steps / (ms_inf / 1e3), 2), "unit": UNIT,
                     "ms_per_step": round(ms_inf / args.steps, 3), "batch_per_gpu": B, "bn_mode": "eval",
                     "pckh_total_joints": int(res["total"][:, 0].sum().item())}
        m.nStack, m.nOutChannels = NSTACK, NJOINT

    # ---- other configurations of BASELINE.json (1 GPU only): training steps of the no-max-pool / strided-block family
    # (configs[3], try_with_aspp_remove_max_pool.creatModel: 3 stacks, heads 2/20/17, Q4 blocks with stride-2 3x3 and
    # 1x1 convolutions -- all on the tensor-core kernels) and of the skeleton+keypoint family (configs[2]) ----------
    extras = None
    if world == 1 and not args.no_extras:
        import importlib

        extras = {}
        for key, modname, note in (("c4_no_max_pool_train", "try_with_aspp_remove_max_pool",
                                    "BASELINE configs[3]: 3-stack, stride-2 Q4 blocks instead of max-pool"),
                                   ("c3_skeleton_keypoints_train", "try_skeleton_and_keypoints",
                                    "BASELINE configs[2]: 4-stack, 38-channel head with the folded limb mix")):
            try:
                fm = importlib.import_module("progressive_process_for_human_pose_estimation_b200." + modname)
                torch.manual_seed(0)
                fnet = fm.creatModel().to(dev)
                fopt = hg.Adam(fnet.parameters(), lr=1e-4)
                with torch.no_grad():
                    shapes = [tuple(o.shape) for o in fnet(x)]
                tg = [torch.rand(sh, device=dev) for sh in shapes]

                def fstep():
                    out = fnet(x)
                    loss = hg.mse_losses(out, tg[0]).sum() if all(sh == shapes[0] for sh in shapes) else \
                        sum(torch.nn.functional.mse_loss(o, t) for o, t in zip(out, tg))
                    fopt.zero_grad()
                    loss.backward()
                    fopt.step()

                for _ in range(3):
                    fstep()
                ms_f = timed(fstep, args.steps)
                fplan = list(fnet.__dict__["_plans"].values())[-1]
                names = [c.name for c in fplan.fwd_calls + fplan.bwd_calls]
                extras[key] = {"value": round(B * args.steps / (ms_f / 1e3), 2), "unit": UNIT,
                               "ms_per_step": round(ms_f / args.steps, 3), "batch_per_gpu": B, "note": note,
                               "loss": "MSE against uniform random targets on every output (throughput only)",
                               "conv_calls": sum(1 for n_ in names if n_.startswith("hg_conv_")),
                               "launches_per_step": len(names)}
                del fnet, fopt, tg
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                extras[key] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    # ---- per-kernel profile of one step (eager, CUDA events around every C-ABI call) ------------------
    roofline, table = None, None
    if rank == 0:
        plan = list(net.__dict__["_plans"].values())[-1]
        plan.profile_records = []
        saved_reducer, plan.reducer = plan.reducer, None
        torch.cuda._sleep(int(1.5e9))  # let the host run ahead so kernels execute back to back
        out = net(x)
        lp = mse[0](out[0], y)
        for k in range(1, NSTACK):
            lp = lp + mse[k](out[k], y)
        opt.zero_grad()
        lp.backward()
        torch.cuda.synchronize()
        recs, plan.profile_records, plan.reducer = plan.profile_records, None, saved_reducer
        agg = {}
        for name, tag, e0, e1 in recs:
            k = (name, tag)
            t = e0.elapsed_time(e1)
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += t
        total_ms = sum(v[1] for v in agg.values())
        table = sorted(((k[0], k[1], v[0], v[1]) for k, v in agg.items()), key=lambda r: -r[3])
        peaks, src = measured_peaks()
        traffic_db = {}
        for tp in ("r02g_roofline_traffic.json", "r02f_roofline_traffic.json", "r02_roofline_traffic.json", "r01_roofline_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, B=32)
                traffic_db = json.load(open(tpath))
                break
        # Every timed entry point that has a roofline: algorithmic FLOPs (conv) or bytes (BatchNorm streaming passes)
        # per launch / its average CUDA-event duration in this step.  Kernels are timed one launch at a time on an
        # otherwise idle GPU at full clocks, so the BURST peak is the divisor (MEASURED_PEAKS.json bf16_tflops / hbm_gbs).
        rows = [r for r in (kernel_roofline(name, tag, n, t, B, peaks, src, total_ms, traffic_db)
                            for name, tag, n, t in table) if r is not None]
        if rows:
            roofline = dict(rows[0])   # the table is sorted by total time: the first row IS the dominant kernel
            roofline["other_kernels"] = rows[1:8]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "kernel_table.txt"), "w") as f:
            f.write(f"# per-call CUDA-event times of one eager step, B={B}/GPU, total {total_ms:.3f} ms\n")
            f.write("# entry point | shape | calls | total ms | avg us | share\n")
            for name, tag, n, t in table:
                f.write(f"{name:24s} {tag:28s} {n:5d} {t:9.3f} {t / n * 1e3:9.2f} {t / total_ms:7.4f}\n")

    # ---- CPU baseline: the oracle port of the reference on the host cores (bounded sample) ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(steps=2, warmup=1, B=2)

    if rank == 0:
        ms_step = ms / args.steps
        pk = measured_peaks()[0]
        model_tf = value / world * TRAIN_GFLOP_PER_IMG / 1e3
        if roofline is not None:
            # the whole step against the tensor roofline: conv FLOPs of the network (291.51 GFLOP / image)
            # / step time, divided by the SUSTAINED peak (a kernel timed inside a long step) and by the burst peak
            roofline["step_model_tflops"] = round(model_tf, 1)
            roofline["step_frac"] = round(model_tf / float(pk["bf16_tflops_sustained"]), 4)
            roofline["step_frac_of_burst"] = round(model_tf / float(pk["bf16_tflops"]), 4)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "8-stack weight-shared hourglass (try_with_torch.creatModel, nStack=8), 16 joints, "
                                   "256x256, fwd + 8x MSE + bwd + Adam", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "l2": "inputs >> L2: ~35 GB of activations touched per step"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches_per_step) * args.steps,
            "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "model_tflops": round(value / world * TRAIN_GFLOP_PER_IMG / 1e3, 1),
            "frac_of_bf16_sustained_peak": round(value / world * TRAIN_GFLOP_PER_IMG / 1e3
                                                 / float(measured_peaks()[0]["bf16_tflops_sustained"]), 4),
            "loss_after_warmup": loss0,
            "phases": phases,
            "loss_and_optimizer": "torch (nn.MSELoss x8, torch.optim.Adam)" if stock else "hg.mse_losses + hg.Adam",
            "e2e_u8_input": e2e_u8,
            "inference": inference,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_reference(steps, warmup, B):
    """The reference's own CPU implementation of the path -- restated by oracle/hourglass_torch.py, pinned
    bit-for-bit to the reference classes (tests/test_oracle_model.py) -- timed on the host cores: forward of the
    8-stack network, eight MSE losses, backward, Adam, on a bounded sample of B images per step."""
    import progressive_process_for_human_pose_estimation_b200.try_with_torch as m
    from oracle import hourglass_torch as ho

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m.nStack, m.nOutChannels = NSTACK, NJOINT
    torch.manual_seed(0)
    sd = ho.clone_state(m.creatModel().state_dict(), requires_grad=True)
    params = [v for k, v in sd.items() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(B, 3, IMG, IMG, generator=g)
    y = torch.rand(B, NJOINT, IMG // 4, IMG // 4, generator=g)
    cfg = ho.Config(nStack=NSTACK, nOutChannels=NJOINT)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = ho.creat_model_s(sd, x, cfg)
        total, _ = ho.mse_losses(out, y)
        opt.zero_grad()
        total.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return {"value": round(B / med, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} timed steps (+{warmup} warm-up) of the same 8-stack train step on {B} images/step, "
                      f"fp32, torch CPU {torch.__version__}, median step {med:.2f} s"}


def run_reference(args):
    """--impl reference: rank 0 times the CPU implementation; other ranks exit without work."""
    world, rank, _ = dist_info()
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    cpu = cpu_reference(steps=steps, warmup=warmup, B=2)
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(2 / cpu["value"] * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8-stack weight-shared hourglass (try_with_torch.creatModel, nStack=8), 16 joints, "
                                   "256x256, fwd + 8x MSE + bwd + Adam", "batch_per_gpu": 32,
                       "parallelism": "cpu", "note": "CPU arm: each step is a bounded sample of 2 images"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_torch_eager(args):
    """--impl torch-eager (informational; NOT the reference arm): the oracle port of the reference network -- plain
    torch.nn.functional calls, i.e. what the unmodified reference modules dispatch -- executed by PyTorch eager on
    ONE B200 (cuDNN / ATen kernels: the "existing Blackwell kernels" of SURVEY 8d, try_with_torch.py:329-344), same
    workload as the product arm (8 stacks, 16 joints, B images, fwd + 8x MSE + bwd + Adam), CUDA-event timed, in fp32
    (TF32 off, as torch defaults for convolutions are ON: both reported) and under torch.autocast(bfloat16) with
    channels_last inputs.  Prints one JSON line with "impl": "torch-eager"."""
    world, rank, local = dist_info()
    if rank != 0:
        return
    import progressive_process_for_human_pose_estimation_b200.try_with_torch as m
    from oracle import hourglass_torch as ho

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    m.nStack, m.nOutChannels = NSTACK, NJOINT
    torch.manual_seed(0)
    sd_cpu = m.creatModel().state_dict()
    g = torch.Generator().manual_seed(100)
    x = torch.randn(B, 3, IMG, IMG, generator=g).to(dev)
    y = torch.rand(B, NJOINT, IMG // 4, IMG // 4, generator=g).to(dev)
    cfg = ho.Config(nStack=NSTACK, nOutChannels=NJOINT)
    warmup, steps = max(3, args.warmup), args.steps
    results = {}
    for mode in ("fp32", "tf32", "bf16_autocast_channels_last"):
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        sd = {k: v.detach().clone().to(dev) for k, v in sd_cpu.items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
        params = [v for v in sd.values() if v.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4)
        xb = x.contiguous(memory_format=torch.channels_last) if mode.startswith("bf16") else x

        def step():
            if mode.startswith("bf16"):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = ho.creat_model_s(sd, xb, cfg)
                out = [o.float() for o in out]
            else:
                out = ho.creat_model_s(sd, xb, cfg)
            total, _ = ho.mse_losses(out, y)
            opt.zero_grad()
            total.backward()
            opt.step()
            return total

        try:
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            results[mode] = {"images_per_s": round(B / (ms / 1e3), 2), "ms_per_step": round(ms, 3),
                             "model_tflops": round(B / (ms / 1e3) * TRAIN_GFLOP_PER_IMG / 1e3, 1),
                             "loss": float(loss)}
        except Exception as e:  # noqa: BLE001  (e.g. out of memory): report, keep going
            results[mode] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        del sd, params, opt
        torch.cuda.empty_cache()
    best = max((r.get("images_per_s", 0.0) for r in results.values()), default=0.0)
    line = {"impl": "torch-eager", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": warmup, "higher_is_better": True, "data": "synthetic", "dtype": "fp32 / tf32 / bf16 autocast",
            "config": {"workload": "oracle port of try_with_torch.creatModel (nStack=8, 16 joints, 256x256) on PyTorch "
                                   f"eager {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, "
                                   "fwd + 8x MSE + bwd + torch.optim.Adam", "batch_per_gpu": B},
            "modes": results, "note": "informational baseline: the library kernels the reference itself would run on "
                                      "this GPU; `value` = the fastest mode"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2] / configs[3] training-step extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
