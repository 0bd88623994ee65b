"""ResNet-18/50 forward with explicit-mask MC-dropout -- CPU oracle.  TEST INFRASTRUCTURE ONLY.

The reference has no classifier; requirements.txt:2 (torchvision) is the only
pointer.  This restates stock ``torchvision.models.resnet18/resnet50`` (eval mode, BN folded)
as a functional forward so that (a) dropout masks come from the Philox contract instead of
torch's CPU generator and (b) an optional bf16-emulation mode rounds exactly where the CUDA
path rounds (weights, stored activations), leaving only fp32 accumulation order as a
difference.  ``forward(..., emulate_bf16=False)`` with T=1 is pinned against torchvision's
own forward in tests/test_oracle.py.

MC-dropout spec: elementwise dropout with probability p on the output of
every residual block (after the final ReLU) and on the pooled feature before fc.  Mask lane
for NHWC offset e of an activation: one BYTE of Philox(c0=e//16, c1=global image, c2=t,
c3=stream(DROPOUT, layer_id)) -- channel c = e % 16 of the chunk reads byte c of the call's
16 output bytes (x0's low byte first).  p is quantised to thr8 / 256 with thr8 = round(p * 256): a value is
dropped iff its byte < thr8 and kept values are scaled by the exact inverse of the realised
keep probability, fl32(256 / (256 - thr8)) -- i.e. torch dropout at p_q = thr8 / 256.
T == 1 disables dropout (deterministic MSP path).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import philox as px

FC_LAYER_ID = 255


def build_torchvision(model="resnet18", num_classes=10, weights_seed=0, logit_gain=None):
    import torchvision
    torch.manual_seed(weights_seed)
    net = getattr(torchvision.models, model)(weights=None, num_classes=num_classes).eval()
    if logit_gain is not None:
        apply_logit_gain(net, logit_gain, weights_seed)
    return net


def apply_logit_gain(net, gain, seed=0):
    """Documented fixture (SURVEY.md section 7 hard part 4): random-init ResNets give degenerate
    confidences, so scale fc.weight by `gain` and add a seeded N(0,1) bias so that confidences
    span the ECE bins and the failure threshold tau is exercised."""
    g = torch.Generator().manual_seed(1000 + seed)
    with torch.no_grad():
        net.fc.weight.mul_(gain)
        net.fc.bias.copy_(torch.randn(net.fc.bias.shape, generator=g))
    return net


def _fold(conv_w, bn, prefix, sd):
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    s = g / torch.sqrt(v + bn.eps)
    return conv_w * s[:, None, None, None], b - m * s


def fold_resnet(net):
    """-> list of conv dicts in execution order + fc.  All fp32 torch tensors (OIHW)."""
    sd = net.state_dict()
    convs = []

    def add(name, conv, bn, bn_name):
        w, b = _fold(sd[name + ".weight"], bn, bn_name, sd)
        convs.append(dict(name=name, w=w.float(), b=b.float(), stride=conv.stride[0], pad=conv.padding[0]))

    add("conv1", net.conv1, net.bn1, "bn1")
    blocks = []
    for li in range(1, 5):
        layer = getattr(net, f"layer{li}")
        for bi, blk in enumerate(layer):
            p = f"layer{li}.{bi}"
            names = ["conv1", "conv2"] + (["conv3"] if hasattr(blk, "conv3") else [])
            idx = []
            for k, cn in enumerate(names):
                add(f"{p}.{cn}", getattr(blk, cn), getattr(blk, f"bn{k + 1}"), f"{p}.bn{k + 1}")
                idx.append(len(convs) - 1)
            ds = None
            if blk.downsample is not None:
                add(f"{p}.downsample.0", blk.downsample[0], blk.downsample[1], f"{p}.downsample.1")
                ds = len(convs) - 1
            blocks.append(dict(convs=idx, ds=ds))
    return dict(convs=convs, blocks=blocks, fc_w=sd["fc.weight"].float(), fc_b=sd["fc.bias"].float())


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def dropout_threshold(p):
    """thr8 = round(p * 256), clipped to [0, 255]."""
    return int(min(max(np.floor(float(p) * 256.0 + 0.5), 0.0), 255.0))


def dropout_scale(p):
    return np.float32(256.0) / np.float32(256 - dropout_threshold(p))


# channel c (0..15) of a 16-channel chunk -> (word, byte) of the chunk's Philox call
_DROP_WORD = np.array([c >> 2 for c in range(16)])
_DROP_BYTE = np.array([c & 3 for c in range(16)])


def dropout_mask(n_images, first_image, t, layer_id, elems_per_image, p, seed):
    """float32 [n_images, elems_per_image]: 0 where dropped else fl32(256 / (256 - thr8))."""
    assert elems_per_image % 16 == 0
    img = np.arange(first_image, first_image + n_images, dtype=np.uint64)[:, None]
    ch = np.arange(elems_per_image // 16, dtype=np.uint64)[None, :]
    xs = px.philox4x32_10(ch, img, t, px.stream_id(px.KIND_DROPOUT, layer_id), seed)
    words = np.stack([np.asarray(x, dtype=np.uint32) for x in xs], axis=-1)                  # [n, chunks, 4]
    lanes = (words[..., _DROP_WORD] >> (8 * _DROP_BYTE).astype(np.uint32)) & np.uint32(0xFF)   # [n, chunks, 16]
    keep = lanes.reshape(n_images, elems_per_image) >= np.uint32(dropout_threshold(p))
    return np.where(keep, dropout_scale(p), np.float32(0.0)).astype(np.float32)


def _drop(x_nchw, t, layer_id, p, seed, first_image):
    """x NCHW fp32 torch; the mask is indexed by NHWC offset (the device layout)."""
    n, c, h, w = x_nchw.shape
    m = dropout_mask(n, first_image, t, layer_id, h * w * c, p, seed).reshape(n, h, w, c)
    return x_nchw * torch.from_numpy(m).permute(0, 3, 1, 2)


@torch.no_grad()
def forward(folded, x_nhwc, T=1, p=0.2, seed=0, first_image=0, emulate_bf16=False):
    """x_nhwc: float32 numpy [N,H,W,3] (already corrupted + normalised).
    Returns logits float32 numpy [N, T, C]."""
    q = _bf16 if emulate_bf16 else (lambda t: t)
    x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).permute(0, 3, 1, 2).contiguous().float()
    x = q(x)
    convs = folded["convs"]

    def conv(i, inp, relu, res=None):
        c = convs[i]
        y = F.conv2d(inp, q(c["w"]), None, stride=c["stride"], padding=c["pad"]) + c["b"][None, :, None, None]
        if res is not None:
            y = y + res
        return torch.relu(y) if relu else y

    stem = q(conv(0, x, True))
    stem = F.max_pool2d(stem, 3, 2, 1)
    outs = []
    use_drop = T > 1
    cache = {}                      # pass-invariant prefix (everything before the first mask)
    for t in range(T):
        h = stem
        for bi, blk in enumerate(folded["blocks"]):
            if bi == 0 and "b0" in cache:
                y = cache["b0"]
            else:
                # device: the 1x1 downsample branch is fused into the last conv as extra K-blocks of the same fp32
                # accumulator, so its result is never rounded to bf16 on its own
                ident = h if blk["ds"] is None else conv(blk["ds"], h, False)
                y = h
                for k, ci in enumerate(blk["convs"]):
                    last = k == len(blk["convs"]) - 1
                    y = conv(ci, y, True, ident if last else None)
                    if not last:
                        y = q(y)
                if bi == 0:
                    cache["b0"] = y
            if use_drop:
                y = _drop(y, t, bi, p, seed, first_image)
            h = q(y)
        feat = h.mean(dim=(2, 3), keepdim=True)
        if use_drop:
            feat = _drop(feat, t, FC_LAYER_ID, p, seed, first_image)
        feat = q(feat).flatten(1)
        outs.append(feat @ q(folded["fc_w"]).t() + folded["fc_b"][None, :])
    return torch.stack(outs, dim=1).numpy().astype(np.float32)


def count_macs(folded, h, w):
    """(prefix MACs, per-pass MACs) per image with the SURVEY.md 8(d) convention
    (prefix = stem only; conv + fc MACs, padded taps counted)."""
    convs, macs = folded["convs"], []
    def out_hw(hh, c):
        k = c["w"].shape[2]
        return (hh + 2 * c["pad"] - k) // c["stride"] + 1
    c0 = convs[0]
    oh, ow = out_hw(h, c0), out_hw(w, c0)
    prefix = oh * ow * c0["w"].numel()
    oh, ow = (oh + 2 - 3) // 2 + 1, (ow + 2 - 3) // 2 + 1
    per = 0
    for blk in folded["blocks"]:
        ih, iw = oh, ow
        for ci in blk["convs"]:
            c = convs[ci]
            oh, ow = out_hw(oh, c), out_hw(ow, c)
            per += oh * ow * c["w"].numel()
        if blk["ds"] is not None:
            c = convs[blk["ds"]]
            per += out_hw(ih, c) * out_hw(iw, c) * c["w"].numel()
    per += folded["fc_w"].numel()
    return prefix, per
