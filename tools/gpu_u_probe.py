"""Localise U-family mismatches: truncated drop-in models vs the torch oracle on the GPU box."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_process_for_human_pose_estimation_b200 as hg
import progressive_process_for_human_pose_estimation_b200.hourglass_compare as m
from progressive_process_for_human_pose_estimation_b200._modules import HGModule
from oracle import hourglass_torch as ho

def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

hg.set_compute_dtype(torch.float32)
torch.manual_seed(0)
net = m.creatModel()
sd = ho.clone_state(net.state_dict())
x = torch.randn(2, 3, 128, 128)

class Trunc(HGModule):
    _is_model = True
    def __init__(self, net, depth):
        super().__init__()
        self.net = net
        self.depth = depth
    def _config_key(self):
        return (self.depth,)
    def _emit(self, b, x):
        pre = self.net.preprocess1
        d = self.depth
        v = b.stem(pre[0], x, relu=False)
        if d >= 1: v = b.bn_relu(pre[1], v)
        if d >= 2: v = pre[3]._emit(b, v)
        if d >= 3: v = b.maxpool2(v)
        if d >= 4: v = pre[5]._emit(b, v)
        if d >= 5: v = pre[6]._emit(b, v)
        inter = v
        if d >= 6:
            st = self.net.stage1
            v = st[0]._emit(b, v)
        if d >= 7: v = st[1]._emit(b, v)
        if d >= 8: v = b.bn_relu(st[3], b.conv(st[2], v))
        if d >= 9:
            t = b.conv(self.net.stage1_out, v, head=True)
            return [t]
        b.output(v)
        return [v]

def oracle(d):
    s = ho.clone_state(sd)
    v = ho._conv(s, "preprocess1.0", x, stride=2, padding=3)
    if d >= 1: v = F.relu(ho._bn(s, "preprocess1.1", v, True))
    if d >= 2: v = ho.residual_block_q4(s, "preprocess1.3", v, 1, True)
    if d >= 3: v = F.max_pool2d(v, 2, 2)
    if d >= 4: v = ho.residual_block_q4(s, "preprocess1.5", v, 1, True)
    if d >= 5: v = ho.residual_block_q4(s, "preprocess1.6", v, 1, True)
    if d >= 6: v = ho.hourglass_u(s, "stage1.0", v, True)
    if d >= 7: v = ho.residual_block_q4(s, "stage1.1", v, 1, True)
    if d >= 8: v = F.relu(ho._bn(s, "stage1.3", ho._conv(s, "stage1.2", v), True))
    if d >= 9: v = ho._conv(s, "stage1_out", v)
    return v

net = net.cuda()
for d in range(10):
    t = Trunc(net, d).cuda()
    with torch.no_grad() if False else torch.enable_grad():
        y = t(x.cuda())[0]
    print("depth", d, tuple(y.shape), rel(y.detach().cpu(), oracle(d)), flush=True)
