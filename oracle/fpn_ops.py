"""Oracle restatement of FPN (ResNet[3,4,5,3] / VGG16 backbones), the Discriminator and the
segmentation losses (rows a1-a3 of SURVEY.md §8), functional over a reference-keyed state dict.
Plain fp32 PyTorch.  Test infrastructure."""
from __future__ import annotations

import torch
import torch.nn.functional as F

RESNET_LAYERS = (3, 4, 5, 3)          # models/fpnseg.py:295  ("ResNet50" builds [3,4,5,3])
VGG_BLOCKS = ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3))   # fpnseg.py:27-142


def _bn(x, p, name, training):
    return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"],
                        p[name + ".bias"], training, 0.1, 1e-5)


def _bottleneck(x, p, pre, stride, training):
    """Bottleneck.forward (fpnseg.py:192-212); stride sits on conv2 (:184)."""
    out = torch.relu(_bn(F.conv2d(x, p[pre + "conv1.weight"]), p, pre + "bn1", training))
    out = torch.relu(_bn(F.conv2d(out, p[pre + "conv2.weight"], stride=stride, padding=1), p, pre + "bn2", training))
    out = _bn(F.conv2d(out, p[pre + "conv3.weight"]), p, pre + "bn3", training)
    if (pre + "downsample.0.weight") in p:
        x = _bn(F.conv2d(x, p[pre + "downsample.0.weight"], stride=stride), p, pre + "downsample.1", training)
    return torch.relu(out + x)


def resnet_features(x, p, pre="back_bone.", training=True):
    """ResNet.forward (fpnseg.py:251-266): returns [c1..c5]."""
    x = torch.relu(_bn(F.conv2d(x, p[pre + "conv1.weight"], stride=2, padding=3), p, pre + "bn1", training))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = [x]
    for li, blocks in enumerate(RESNET_LAYERS, start=1):
        for bi in range(blocks):
            stride = 2 if (bi == 0 and li > 1) else 1
            x = _bottleneck(x, p, f"{pre}layer{li}.{bi}.", stride, training)
        feats.append(x)
    return feats


def vgg16_features(x, p, pre="back_bone.", training=True):
    """VGG16.forward (fpnseg.py:154-166): five (conv3x3+BN+ReLU)*k + maxpool blocks."""
    feats = []
    for bi, (_, nconv) in enumerate(VGG_BLOCKS, start=1):
        for ci in range(nconv):
            base = f"{pre}block_{bi}.{3 * ci}"
            x = F.conv2d(x, p[base + ".weight"], p[base + ".bias"], padding=1)
            x = torch.relu(_bn(x, p, f"{pre}block_{bi}.{3 * ci + 1}", training))
        x = F.max_pool2d(x, 2, 2)
        feats.append(x)
    return feats


def _up(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)   # fpnseg.py:358-359


def _conv(x, p, name, padding=0):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], padding=padding)


def _gn(x, p, name):
    c = x.shape[1]
    return F.group_norm(x, c, p[name + ".weight"], p[name + ".bias"], 1e-5)     # GroupNorm(C, C), fpnseg.py:354-355


def fpn_forward(x, p, backbone="resnet", training=True):
    """FPN.forward (fpnseg.py:391-444) -> (logits, [p2,p3,p4,p5])."""
    feats = resnet_features(x, p, "back_bone.", training) if backbone == "resnet" \
        else vgg16_features(x, p, "back_bone.", training)
    c2, c3, c4, c5 = feats[1:]
    p5 = _conv(c5, p, "toplayer")
    l4 = _conv(c4, p, "latlayer1")
    p4 = _up(p5, l4.shape[-2:]) + l4                                             # _upsample_add :371-388
    l3 = _conv(c3, p, "latlayer2")
    p3 = _up(p4, l3.shape[-2:]) + l3
    l2 = _conv(c2, p, "latlayer3")
    p2 = _up(p3, l2.shape[-2:]) + l2
    fmap = [p2, p3, p4, p5]                                                       # pre-smoothing :415-418
    s4_in = _conv(p4, p, "smooth1", 1)
    s3_in = _conv(p3, p, "smooth2", 1)
    s2_in = _conv(p2, p, "smooth3", 1)
    hw = s2_in.shape[-2:]

    def block256(t):   # conv2 -> gn2 -> relu -> upsample
        return _up(torch.relu(_gn(_conv(t, p, "conv2", 1), p, "gn2")), hw)

    def block128(t, up=True):
        y = torch.relu(_gn(_conv(t, p, "semantic_branch", 1), p, "gn1"))
        return _up(y, hw) if up else y

    s5 = block128(block256(block256(p5)))                                         # :428-432
    s4 = block128(block256(s4_in))                                                # :435-437
    s3 = block128(s3_in)                                                          # :440
    s2 = block128(s2_in, up=False)                                                # :442
    logits = _up(_conv(s2 + s3 + s4 + s5, p, "conv3"), (4 * hw[0], 4 * hw[1]))    # :444
    return logits, fmap


def discriminator_loss(feat_s, feat_t, p, lambda_=0.02):
    """Discriminator.forward (fpnseg.py:496-511) with GradientReversalFunction
    (gradient_reversal.py:6-24): 4 x (conv3x3 + GN32 + ReLU) + conv3x3 -> BCE vs 1 (source) / 0 (target)."""
    def tower(t):
        t = _GRL.apply(t, lambda_)
        for i in range(4):
            t = F.conv2d(t, p[f"dis_tower.{3 * i}.weight"], p[f"dis_tower.{3 * i}.bias"], padding=1)
            t = torch.relu(F.group_norm(t, 32, p[f"dis_tower.{3 * i + 1}.weight"], p[f"dis_tower.{3 * i + 1}.bias"], 1e-5))
        return F.conv2d(t, p["cls_logits.weight"], p["cls_logits.bias"], padding=1)

    xs, xt = tower(feat_s), tower(feat_t)
    return F.binary_cross_entropy_with_logits(xs, torch.ones_like(xs)) + \
        F.binary_cross_entropy_with_logits(xt, torch.zeros_like(xt))


class _GRL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = lambda_
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return -ctx.lambda_ * g, None


def grad_reverse(x, lambda_):
    return _GRL.apply(x, lambda_)


def dice_loss(logits, target, smooth=1.0, pw=2):
    """DiceLoss(BinaryDiceLoss) (utils/losses.py:43-95): softmax over channels, per-channel dice."""
    prob = F.softmax(logits, dim=1)
    total = 0
    for c in range(target.shape[1]):
        pr = prob[:, c].contiguous().view(prob.shape[0], -1)
        tg = target[:, c].contiguous().view(target.shape[0], -1)
        num = (pr * tg).sum(1) + smooth
        den = (pr.pow(pw) + tg.pow(pw)).sum(1) + smooth
        total = total + (1 - num / den).mean()
    return total / target.shape[1]


def seg_loss(logits, target):
    """dice + BCE-with-logits, as train_cardiac_uda.py:228."""
    return dice_loss(logits, target) + F.binary_cross_entropy_with_logits(logits, target)


def overlap_metrics(target, pred):
    """_calculate_overlap_metrics (train_cardiac_uda.py:496-511): pixel-acc, dice, precision,
    specificity, recall from binary masks."""
    eps = 1e-5
    out, gt = pred.reshape(-1).float(), target.reshape(-1).float()
    tp = (out * gt).sum()
    fp = (out * (1 - gt)).sum()
    fn = ((1 - out) * gt).sum()
    tn = ((1 - out) * (1 - gt)).sum()
    pixel_acc = (tp + tn + eps) / (tp + tn + fp + fn + eps)
    dice = (2 * tp + eps) / (2 * tp + fp + fn + eps)
    precision = (tp + eps) / (tp + fp + eps)
    specificity = (tn + eps) / (tn + fp + eps)
    recall = (tp + eps) / (tp + fn + eps)
    return pixel_acc, dice, precision, specificity, recall
