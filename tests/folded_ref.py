"""CPU evaluation of the network from the FOLDED tensors the packer emits (test helper).

It follows the data flow of the CUDA engine (bias+act fused convs, per-level folded head weights,
pre-normalised fusion scalars) with torch ops, so a wrong fold / name / layout in
hmd_ego_pose_b200/packer.py shows up on the CPU, before any GPU run."""
import math

import torch
import torch.nn.functional as F

from hmd_ego_pose_b200 import packer


def _pad(x, k, s):
    h = x.shape[-1]
    extra = (math.ceil(h / s) - 1) * s - h + k
    lo = extra // 2
    return F.pad(x, [lo, extra - lo, lo, extra - lo])


def swish(x):
    return x * torch.sigmoid(x)


def forward(folded, x, num_classes=1):
    t = {k: torch.from_numpy(v) for k, v in folded.items()}
    pw = lambda x, n: F.conv2d(x, t[n + ".w"][:, :, None, None], t[n + ".b"])
    dw = lambda x, n, k, s, bias=None: F.conv2d(_pad(x, k, s), t[n][:, None], bias, stride=s, groups=x.shape[1])
    pool = lambda x: F.max_pool2d(_pad(x, 3, 2), 3, 2)
    up = lambda x: F.interpolate(x, scale_factor=2, mode="nearest")
    x = swish(F.conv2d(_pad(x, 3, 2), t["stem.w"], t["stem.b"], stride=2))
    outs = []
    for i, (k, s, e, cin, cout, skip) in enumerate(packer.B0_BLOCKS):
        q, inp = f"blk{i}", x
        if e != 1:
            x = swish(pw(x, q + ".exp"))
        x = swish(dw(x, q + ".dw.w", k, s, t[q + ".dw.b"]))
        g = x.mean((2, 3))
        g = swish(g @ t[q + ".se_r.w"].T + t[q + ".se_r.b"])
        g = torch.sigmoid(g @ t[q + ".se_e.w"].T + t[q + ".se_e.b"])
        x = pw(x * g[:, :, None, None], q + ".proj")
        if skip:
            x = x + inp
        outs.append(x)
    P3, P4, P5 = outs[4], outs[10], outs[15]
    feats = None
    for c in range(3):
        q = f"bifpn{c}"
        node = lambda n, z: pw(dw(z, f"{q}.{n}.dw.w", 3, 1), f"{q}.{n}.pw")
        w = lambda n: t[f"{q}.fw.{n}"]
        if c == 0:
            p3i, p4i, p5i = pw(P3, q + ".p3_dc"), pw(P4, q + ".p4_dc"), pw(P5, q + ".p5_dc")
            p6i = pool(pw(P5, q + ".p5_to_p6"))
            p7i = pool(p6i)
            p4i2, p5i2 = pw(P4, q + ".p4_dc2"), pw(P5, q + ".p5_dc2")
        else:
            p3i, p4i, p5i, p6i, p7i = feats
            p4i2, p5i2 = p4i, p5i
        p6u = node("conv6_up", swish(w("p6_w1")[0] * p6i + w("p6_w1")[1] * up(p7i)))
        p5u = node("conv5_up", swish(w("p5_w1")[0] * p5i + w("p5_w1")[1] * up(p6u)))
        p4u = node("conv4_up", swish(w("p4_w1")[0] * p4i + w("p4_w1")[1] * up(p5u)))
        p3o = node("conv3_up", swish(w("p3_w1")[0] * p3i + w("p3_w1")[1] * up(p4u)))
        p4o = node("conv4_down", swish(w("p4_w2")[0] * p4i2 + w("p4_w2")[1] * p4u + w("p4_w2")[2] * pool(p3o)))
        p5o = node("conv5_down", swish(w("p5_w2")[0] * p5i2 + w("p5_w2")[1] * p5u + w("p5_w2")[2] * pool(p4o)))
        p6o = node("conv6_down", swish(w("p6_w2")[0] * p6i + w("p6_w2")[1] * p6u + w("p6_w2")[2] * pool(p5o)))
        p7o = node("conv7_down", swish(w("p7_w2")[0] * p7i + w("p7_w2")[1] * pool(p6o)))
        feats = (p3o, p4o, p5o, p6o, p7o)
    rows = lambda z, wd: z.permute(0, 2, 3, 1).contiguous().view(z.shape[0], -1, wd)
    res = {k: [] for k in ("box", "cls", "rot", "trans", "hand")}
    widths = {"box": 4, "cls": num_classes, "rot": 3, "hand": 63}
    for lvl, f in enumerate(feats):
        for h in res:
            z = f
            for i in range(3):
                z = swish(pw(dw(z, f"head.{h}.l{i}.dw.w", 3, 1), f"head.{h}.l{i}.lvl{lvl}.pw"))
            if h == "trans":
                xy = rows(pw(dw(z, "head.trans.hdr0.dw.w", 3, 1), "head.trans.hdr0.pw"), 2)
                zz = rows(pw(dw(z, "head.trans.hdr1.dw.w", 3, 1), "head.trans.hdr1.pw"), 1)
                res[h].append(torch.cat((xy, zz), 2))
            else:
                res[h].append(rows(pw(dw(z, f"head.{h}.hdr0.dw.w", 3, 1), f"head.{h}.hdr0.pw"), widths[h]))
    cat = {k: torch.cat(v, 1) for k, v in res.items()}
    return feats, cat["box"], torch.sigmoid(cat["cls"]), cat["rot"], cat["trans"], cat["hand"]
