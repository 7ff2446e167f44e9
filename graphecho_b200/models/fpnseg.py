"""Mirror of models/fpnseg.py: VGG16 / Bottleneck / ResNet / ResNet50 / ResNet101 / FPN / Discriminator
with the reference's constructor signatures, forward contracts and state_dict keys.

B200 design: feature maps are channels_last (NHWC) end to end so the dense convolutions run on the
tensor-core library path without layout transposes (bf16 under torch.autocast, fp32 otherwise); the
full-tensor passes the reference runs between convolutions in the head -- bilinear upsample + add,
GroupNorm(C,C) + ReLU + upsample, branch sum + conv3 + x4 upsample -- are the fused sm_100a kernels in
csrc/fpn_head.cu (reference: fpnseg.py:371-388, 409-444).  Same-size `_upsample` calls
(3 of the 7 in the head) are exact identities and cost nothing here."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as GF
from .gradient_reversal import GradientReversal

__all__ = ["ResNet", "ResNet50", "ResNet101", "VGG16", "FPN", "Discriminator"]

_VGG_PLAN = ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3))     # (width, convs) per block


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("graphecho_b200 modules run on CUDA tensors only (no CPU fallback)")


class VGG16(nn.Module):
    """13 x (conv3x3 + BN + ReLU) in five max-pooled blocks; returns the five pooled maps
    (fpnseg.py:18-166).  Sequential indices match the reference: conv at 3i, BN at 3i+1."""

    def __init__(self, in_channels):
        super().__init__()
        cin = in_channels
        for b, (width, n) in enumerate(_VGG_PLAN, start=1):
            layers = []
            for _ in range(n):
                layers += [nn.Conv2d(cin, width, 3, 1, 1), nn.BatchNorm2d(width), nn.ReLU()]
                cin = width
            layers.append(nn.MaxPool2d(2, 2))
            setattr(self, f"block_{b}", nn.Sequential(*layers))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="leaky_relu")
                if m.bias is not None:
                    m.bias.detach().zero_()

    def _blocks(self, x, first, last):
        feats = []
        for b in range(first, last + 1):
            mods = list(getattr(self, f"block_{b}"))
            for i in range(0, len(mods) - 1, 3):          # (conv, BN, ReLU) triples -> conv + fused BN+ReLU
                x = GF.conv_bn_act(x, mods[i], mods[i + 1], relu=True)    # the conv bias cancels in train-mode BN
            x = mods[-1](x)                                # max-pool
            feats.append(x)
        return feats

    def forward_lower(self, x):
        """blocks 1-4 -> [c1, c2, c3, c4]"""
        return self._blocks(x, 1, 4)

    def forward_upper(self, c4):
        """block 5 -> c5"""
        return self._blocks(c4, 5, 5)[0]

    def forward(self, x):
        feats = self.forward_lower(x)
        return feats + [self.forward_upper(feats[-1])]


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=False)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        # 1x1 convs: tcgen05 GEMM whose epilogue yields the BatchNorm statistics (bf16 path), else cuDNN + statistics pass
        out = GF.conv1x1_bn_act(x, self.conv1, self.bn1, relu=True)
        out = GF.bn_act(self.conv2(out), self.bn2, relu=True)
        identity = x if self.downsample is None else GF.bn_act(self.downsample[0](x), self.downsample[1], relu=False)
        return GF.conv1x1_bn_act(out, self.conv3, self.bn3, residual=identity, relu=True)   # BN + add + ReLU in one pass


class ResNet(nn.Module):
    """Stem (conv7x7/2 + BN + ReLU + maxpool) + 4 stages; returns [c1..c5] (fpnseg.py:214-266)."""

    def __init__(self, block, layers, in_channel, pretrained=False):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channel, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=False)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self._init_weights()
        if pretrained:
            raise RuntimeError("pretrained ImageNet weights are not available offline")

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
        stage = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        stage += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*stage)

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward_lower(self, x):
        """stem + layer1-3 -> [c1, c2, c3, c4]"""
        x = GF.bn_act(GF.stem_conv(x, self.conv1), self.bn1, relu=True)
        x = GF.maxpool3s2(x) if x.shape[1] % 8 == 0 else self.maxpool(x)
        feats = [x]
        for stage in (self.layer1, self.layer2, self.layer3):
            x = stage(x)
            feats.append(x)
        return feats

    def forward_upper(self, c4):
        """layer4 -> c5"""
        return self.layer4(c4)

    def forward(self, x):
        feats = self.forward_lower(x)
        return feats + [self.forward_upper(feats[-1])]


def ResNet50(in_channel=3, pretrained=True):
    """The reference's "ResNet-50" builds stages [3, 4, 5, 3] (fpnseg.py:295); kept for state_dict parity."""
    if pretrained:
        raise RuntimeError("pretrained ImageNet weights are not available offline")
    return ResNet(Bottleneck, [3, 4, 5, 3], in_channel=in_channel)


def ResNet101(in_channel=3, pretrained=True):
    return ResNet(Bottleneck, [3, 4, 23, 3], in_channel=in_channel, pretrained=pretrained)


class FPN(nn.Module):
    """FPN(num_blocks, num_classes, in_channel, back_bone='resnet'|'VGG16', pretrained=False)
    forward(x) -> (logits [B,nc,H,W] fp32, [p2,p3,p4,p5] pre-smoothing)   (fpnseg.py:309-444).
    `num_blocks` is accepted and ignored, as in the reference."""

    def __init__(self, num_blocks, num_classes, in_channel, back_bone="resnet", pretrained=False):
        super().__init__()
        self.in_planes = 64
        self.num_classes = num_classes
        if back_bone == "resnet":
            self.back_bone = ResNet50(in_channel=in_channel, pretrained=pretrained)
            widths = (2048, 1024, 512, 256)
        elif back_bone == "VGG16":
            self.back_bone = VGG16(in_channels=in_channel)
            widths = (512, 512, 256, 128)
        else:
            raise ValueError(f"unknown back_bone {back_bone!r}")
        self.toplayer = nn.Conv2d(widths[0], 256, 1)
        self.latlayer1 = nn.Conv2d(widths[1], 256, 1)
        self.latlayer2 = nn.Conv2d(widths[2], 256, 1)
        self.latlayer3 = nn.Conv2d(widths[3], 256, 1)
        self.smooth1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.smooth2 = nn.Conv2d(256, 256, 3, 1, 1)
        self.smooth3 = nn.Conv2d(256, 256, 3, 1, 1)
        self.semantic_branch = nn.Conv2d(256, 128, 3, 1, 1)
        self.conv2 = nn.Conv2d(256, 256, 3, 1, 1)
        self.conv3 = nn.Conv2d(128, num_classes, 1)
        self.gn1 = nn.GroupNorm(128, 128)
        self.gn2 = nn.GroupNorm(256, 256)

    def _upsample(self, x, h, w):
        return GF.upsample_bilinear(x, (h, w))

    def _upsample_add(self, x, y):
        return GF.upsample_add(x, y)

    def forward_trunk_lower(self, x):
        """First part of forward_trunk: the backbone up to c4 -> (c2, c3, c4).  The engine runs (and captures) the trunk
        in two parts so that the gradients of the upper part -- 2/3 of the backbone's parameters sit in its last stage
        -- can be all-reduced while the lower part's backward is still running."""
        _need_cuda(x)
        x = x.contiguous(memory_format=torch.channels_last)
        _, c2, c3, c4 = self.back_bone.forward_lower(x)
        return c2, c3, c4

    def forward_trunk_upper(self, c2, c3, c4):
        """Second part: last backbone stage + lateral / top-down pyramid -> (p2, p3, p4, p5)."""
        c5 = self.back_bone.forward_upper(c4)
        p5 = GF.conv_bias(c5, self.toplayer)
        p4 = GF.upsample_add(p5, GF.conv_bias(c4, self.latlayer1))
        p3 = GF.upsample_add(p4, GF.conv_bias(c3, self.latlayer2))
        p2 = GF.upsample_add(p3, GF.conv_bias(c2, self.latlayer3))
        return p2, p3, p4, p5

    def upper_trunk_parameters(self):
        """Parameters of forward_trunk_upper (last backbone stage, top and lateral layers)."""
        last = self.back_bone.layer4 if hasattr(self.back_bone, "layer4") else self.back_bone.block_5
        mods = [last, self.toplayer, self.latlayer1, self.latlayer2, self.latlayer3]
        return [p for m in mods for p in m.parameters()]

    def head_parameters(self):
        mods = [self.smooth1, self.smooth2, self.smooth3, self.semantic_branch, self.conv2, self.conv3, self.gn1, self.gn2]
        return [p for m in mods for p in m.parameters()]

    def forward_trunk(self, x):
        """Backbone + lateral / top-down pyramid (fpnseg.py:391-423): x -> [p2, p3, p4, p5] (pre-smoothing,
        what the reference returns as `features_map`)."""
        _need_cuda(x)
        x = x.contiguous(memory_format=torch.channels_last)
        _, c2, c3, c4, c5 = self.back_bone(x)
        p5 = GF.conv_bias(c5, self.toplayer)
        p4 = GF.upsample_add(p5, GF.conv_bias(c4, self.latlayer1))
        p3 = GF.upsample_add(p4, GF.conv_bias(c3, self.latlayer2))
        p2 = GF.upsample_add(p3, GF.conv_bias(c2, self.latlayer3))
        return p2, p3, p4, p5

    def forward_head(self, p2, p3, p4, p5):
        """Smoothing + semantic head (fpnseg.py:424-444): pyramid -> logits."""
        q4, q3, q2 = GF.conv_bias(p4, self.smooth1), GF.conv_bias(p3, self.smooth2), GF.conv_bias(p2, self.smooth3)
        hw = q2.shape[-2:]
        g1, g2 = self.gn1, self.gn2

        # GroupNorm(C, C) normalises every channel on its own, so a per-channel constant added before it
        # cancels exactly: the biases of conv2 / semantic_branch cannot influence the output (their true
        # gradient is identically zero).  Skipping them saves cuDNN's bias-gradient reductions over the
        # largest maps of the head; the parameters stay in the state_dict, untouched.
        def wide(t):        # conv2 -> GroupNorm(256,256) -> ReLU -> _upsample      (fpnseg.py:428-435)
            y = F.conv2d(t, self.conv2.weight, None, 1, 1)
            return GF.gn_relu_upsample(y, g2.weight, g2.bias, hw, g2.eps)

        def narrow(t):      # semantic_branch -> GroupNorm(128,128) -> ReLU -> _upsample
            y = F.conv2d(t, self.semantic_branch.weight, None, 1, 1)
            return GF.gn_relu_upsample(y, g1.weight, g1.bias, hw, g1.eps)

        s5 = narrow(wide(wide(p5)))
        s4 = narrow(wide(q4))
        s3 = narrow(q3)
        s2 = narrow(q2)
        return GF.seg_tail(s2, s3, s4, s5, self.conv3.weight, self.conv3.bias, 4)

    def forward(self, x):
        p2, p3, p4, p5 = self.forward_trunk(x)
        return self.forward_head(p2, p3, p4, p5), [p2, p3, p4, p5]


class Discriminator(nn.Module):
    """Patch domain discriminator on one pyramid level (fpnseg.py:447-511): GRL -> 4 x (conv3x3 +
    GroupNorm(32) + ReLU) -> conv3x3 -> BCE against 1 (source) / 0 (target).  Source and target maps
    go through the tower as one batch (identical maths: GroupNorm is per-sample)."""

    def __init__(self, num_convs=4, in_channels=256, grad_reverse_lambda=-1.0, grl_applied_domain="both",
                 patch_stride=None):
        super().__init__()
        tower = []
        for _ in range(num_convs):
            tower += [nn.Conv2d(in_channels, in_channels, 3, 1, 1), nn.GroupNorm(32, in_channels), nn.ReLU()]
        self.add_module("dis_tower", nn.Sequential(*tower))
        self.cls_logits = nn.Conv2d(in_channels, 1, 3, 1, 1)
        assert patch_stride is None or type(patch_stride) == int, "wrong format of patch stride"
        self.patch_stride = patch_stride
        if self.patch_stride:
            self.pool = nn.AvgPool2d(3, patch_stride, 1)
        for m in list(self.dis_tower.modules()) + [self.cls_logits]:
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.01)
                nn.init.constant_(m.bias, 0)
        self.grad_reverse = GradientReversal(grad_reverse_lambda)
        self.loss_fn = nn.BCEWithLogitsLoss()
        assert grl_applied_domain in ("both", "target")
        self.grl_applied_domain = grl_applied_domain
        self.source_label, self.target_label = 1.0, 0.0

    def _tower(self, x):
        """conv3x3 -> fused GroupNorm(32)+ReLU (one NHWC pass, no fp32 up-cast) per tower stage."""
        x = x.contiguous(memory_format=torch.channels_last)
        mods = list(self.dis_tower)
        for i in range(0, len(mods), 3):
            conv, gn = mods[i], mods[i + 1]
            # conv bias folded into the GroupNorm kernels: no separate bias-add / bias-gradient passes over the map
            y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
            x = GF.gn_relu(y, gn.weight, gn.bias, gn.num_groups, gn.eps, pre_bias=conv.bias)
        return self.cls_logits(x).float()

    def forward_joint(self, feature_all, n_source):
        """forward((feature_all[:n_source], feature_all[n_source:])) without slicing / re-concatenating the
        feature map (the source frames come first): only the small logit map is split."""
        _need_cuda(feature_all)
        x = self._tower(self.grad_reverse(feature_all))
        xs, xt = x[:n_source], x[n_source:]
        return self.loss_fn(xs, torch.full_like(xs, self.source_label)) + \
            self.loss_fn(xt, torch.full_like(xt, self.target_label))

    def forward(self, feature, domain="source"):
        fs, ft = feature
        _need_cuda(fs)
        ns = fs.shape[0]
        if fs.shape[1:] == ft.shape[1:]:
            x = self._tower(self.grad_reverse(torch.cat([fs, ft], dim=0)))
            xs, xt = x[:ns], x[ns:]
        else:
            xs = self._tower(self.grad_reverse(fs))
            xt = self._tower(self.grad_reverse(ft))
        return self.loss_fn(xs, torch.full_like(xs, self.source_label)) + \
            self.loss_fn(xt, torch.full_like(xt, self.target_label))
