"""Fold a torchvision ResNet into the FAVW1 blob consumed by ``fav_load_weights``.

The reference ships no model; ``requirements.txt:2`` (torchvision) is the only pointer, so the
classifier is stock ``torchvision.models.resnet18 / resnet50``.  BatchNorm (eval mode) is folded
into the preceding conv in fp32, weights are laid out [Cout][R][S][Cin] and cast to bf16
(round-to-nearest-even); biases stay fp32.

Blob: "FAVW1\\0\\0\\0", int32 n_convs, num_classes, model_id, n_blocks, then int32[n_blocks][2]
(n_convs_in_block, has_downsample), pad to 16 B, then per conv: int32[8] (cout, cin, r, s,
stride, pad, 0, 0), bf16 weights (padded to 16 B), fp32 bias (padded to 16 B).  Conv order:
stem, then per block conv1, conv2, [conv3], [downsample]; the fc comes last as a 1x1 conv.
"""
import numpy as np
import torch

MODEL_IDS = {"resnet18": 18, "resnet50": 50}


def build_model(model="resnet18", num_classes=10, weights_seed=0, logit_gain=None):
    """Random-init torchvision model (there is no network access for checkpoints)."""
    import torchvision
    torch.manual_seed(weights_seed)
    net = getattr(torchvision.models, model)(weights=None, num_classes=num_classes).eval()
    if logit_gain is not None:
        # documented fixture (hard part 4): spread the confidences of a random-init net
        g = torch.Generator().manual_seed(1000 + weights_seed)
        with torch.no_grad():
            net.fc.weight.mul_(float(logit_gain))
            net.fc.bias.copy_(torch.randn(net.fc.bias.shape, generator=g))
    return net


def _fold(conv, bn):
    w = conv.weight.detach().float()
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * s
    return w * s[:, None, None, None], b


def _pad16(b):
    return b + b"\0" * (-len(b) % 16)


def _record(w_oihw, bias, stride, pad):
    cout, cin, r, s = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)           # [Cout][R][S][Cin]
    hdr = np.array([cout, cin, r, s, stride, pad, 0, 0], dtype=np.int32).tobytes()
    return hdr + _pad16(w.view(torch.int16).numpy().tobytes()) + _pad16(bias.float().numpy().tobytes())


def pack_resnet(net, model="resnet18"):
    """-> bytes (FAVW1).  ``net`` is a torchvision ResNet in eval mode."""
    recs, blocks = [], []
    w, b = _fold(net.conv1, net.bn1)
    recs.append(_record(w, b, net.conv1.stride[0], net.conv1.padding[0]))
    for li in range(1, 5):
        for blk in getattr(net, f"layer{li}"):
            names = ["conv1", "conv2"] + (["conv3"] if hasattr(blk, "conv3") else [])
            for k, cn in enumerate(names):
                conv = getattr(blk, cn)
                w, b = _fold(conv, getattr(blk, f"bn{k + 1}"))
                recs.append(_record(w, b, conv.stride[0], conv.padding[0]))
            if blk.downsample is not None:
                w, b = _fold(blk.downsample[0], blk.downsample[1])
                recs.append(_record(w, b, blk.downsample[0].stride[0], blk.downsample[0].padding[0]))
            blocks.append((len(names), 1 if blk.downsample is not None else 0))
    fcw = net.fc.weight.detach().float()[:, :, None, None]
    recs.append(_record(fcw, net.fc.bias.detach().float(), 1, 0))
    num_classes = net.fc.weight.shape[0]
    head = b"FAVW1\0\0\0" + np.array([len(recs), num_classes, MODEL_IDS[model], len(blocks)], dtype=np.int32).tobytes()
    head += np.array(blocks, dtype=np.int32).tobytes()
    return _pad16(head) + b"".join(recs)
