import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np
import test_families as tf
import progressive_process_for_human_pose_estimation_b200 as hg
hg.set_compute_dtype(torch.bfloat16)
for script, factory, fixture in tf.FAMILIES:
    g = tf.load(fixture, "train")
    worst_l, worst_e = 0, 0
    for rep in range(6):
        net = tf.build(script, factory).cuda()
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        for it in range(3):
            net.load_state_dict(sd0); net.zero_grad(set_to_none=True)
            out, losses = tf.run_step(net, g)
            for i, o in enumerate(out):
                worst_l = max(worst_l, abs(losses[i].item() - g["losses"][i]) / abs(g["losses"][i]))
            worst_e = max(worst_e, tf.rel(out[0].detach().cpu().numpy(), g["out0"]))
    print(f"{script:40s} worst loss dev {worst_l:.3f} (limit 0.1)  worst err0 {worst_e:.3f} (limit {max(0.3, 2*float(g['out_noise_bf16'][0])):.3f})", flush=True)
