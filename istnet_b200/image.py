"""Image branch of IST-Net: ResNet-18 (output stride 8) + PSP head + 3 up-sampling stages + 1x1 head.

Parameter names mirror the reference (`model/resnet.py:37-66,109-214`, `model/modules.py:10-81,234-241`) so that
checkpoints interchange: `model.feats.layer1.0.conv1.weight`, `model.psp.stages.0.1.weight`,
`model.up_1.conv.1.weight`, `model.up_1.conv.3.weight` (PReLU), `model.final.0.weight`, and the never-used
`model.feats.fc.*` (resnet.py:140) which stays in the state dict and never receives a gradient.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    """resnet.py:37-66"""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        res = x if self.downsample is None else self.downsample(x)
        return self.relu(out + res)


class ResNet18(nn.Module):
    """resnet.py:109-202 for BasicBlock x [2,2,2,2].  `_make_layer` (resnet.py:153-180) ignores its dilation
    argument, so layer3/layer4 are stride 1 with dilation 1: total stride 8."""

    def __init__(self, num_classes=1000):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.inplanes = 64
        self.layer1 = self._stage(64, 1)
        self.layer2 = self._stage(128, 2)
        self.layer3 = self._stage(256, 1)
        self.layer4 = self._stage(512, 1)
        self.avgpool = nn.AvgPool2d(7)
        self.fc = nn.Linear(512, num_classes)  # unused on the path, kept for checkpoint compatibility
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _stage(self, planes, stride):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(planes))
        blocks = [BasicBlock(self.inplanes, planes, stride, down), BasicBlock(planes, planes)]
        self.inplanes = planes
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer2(self.layer1(x))
        x_3 = self.layer3(x)
        return self.layer4(x_3), x_3


class PSPModule(nn.Module):
    """modules.py:10-34"""

    def __init__(self, features, out_features=1024, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.stages = nn.ModuleList(
            [nn.Sequential(nn.AdaptiveAvgPool2d(output_size=(s, s)), nn.Conv2d(features, features, kernel_size=1, bias=False)) for s in sizes]
        )
        self.bottleneck = nn.Conv2d(features * (len(sizes) + 1), out_features, kernel_size=1)
        self.relu = nn.ReLU()

    def forward(self, feats):
        h, w = feats.size(2), feats.size(3)
        priors = [F.interpolate(stage(feats), size=(h, w), mode="bilinear", align_corners=False) for stage in self.stages]
        return self.relu(self.bottleneck(torch.cat(priors + [feats], 1)))


class PSPUpsample(nn.Module):
    """modules.py:37-48"""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
            nn.Conv2d(cin, cout, 3, padding=1),
            nn.BatchNorm2d(cout),
            nn.PReLU(),
        )

    def forward(self, x):
        return self.conv(x)


class Modified_PSPNet(nn.Module):
    """modules.py:51-81 for backend resnet18 (psp_size 512); weights are never downloaded (pretrained=False)."""

    def __init__(self, sizes=(1, 2, 3, 6), psp_size=512):
        super().__init__()
        self.feats = ResNet18()
        self.psp = PSPModule(psp_size, 1024, sizes)
        self.drop_1 = nn.Dropout2d(p=0.3)
        self.up_1 = PSPUpsample(1024, 256)
        self.up_2 = PSPUpsample(256, 64)
        self.up_3 = PSPUpsample(64, 64)
        self.drop_2 = nn.Dropout2d(p=0.15)
        self.final = nn.Sequential(nn.Conv2d(64, 128, kernel_size=1), nn.BatchNorm2d(128), nn.PReLU())
        self.dropout_noise_fn = None  # tests may inject (B,C,1,1) masks to compare against a CPU oracle

    def _drop(self, x, p):
        """F.dropout2d == x * bernoulli(1-p)/(1-p) per (b,c) plane (ATen feature_dropout)."""
        if not self.training or p == 0.0:
            return x
        if self.dropout_noise_fn is not None:
            return x * self.dropout_noise_fn(x.shape[0], x.shape[1], p).to(x.device)
        noise = torch.empty(x.shape[0], x.shape[1], 1, 1, device=x.device, dtype=x.dtype).bernoulli_(1 - p).div_(1 - p)
        return x * noise

    def forward(self, x):
        f, _ = self.feats(x)
        p = self._drop(self.psp(f), self.drop_1.p)
        p = self._drop(self.up_1(p), self.drop_2.p)
        p = self._drop(self.up_2(p), self.drop_2.p)
        return self.final(self.up_3(p))


class ModifiedResnet(nn.Module):
    """modules.py:234-241"""

    def __init__(self):
        super().__init__()
        self.model = Modified_PSPNet()

    def forward(self, x):
        """Dense (B,128,H,W) feature map (modules.py:239-241).  On CUDA it runs on the B200 kernels, forward only (the training
        path is gather(): IST_Net reads the map only at the `choose`d pixels); CPU tensors take the plain-torch modules, which
        exist for state-dict / API compatibility and are not part of the product path."""
        if x.is_cuda:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                raise RuntimeError("istnet_b200: ModifiedResnet.forward is forward-only on CUDA; use gather()/gather_rows() for training")
            from .image_engine import dense_map

            return dense_map(self.model, x)
        return self.model(x)

    def gather(self, rgb, choose):
        """rgb (B,3,H,W), choose (B,N) int64 -> per-point pixel features (B,128,N) == `torch.gather(self(rgb).view(b,d,-1), 2, choose)`
        (ist_net.py:41-45), computed channels-last on the B200 kernels (tcgen05 convolutions, fused BN / activation
        passes, the head evaluated only at the chosen pixels).  There is no CPU path."""
        if not rgb.is_cuda:
            raise RuntimeError("istnet_b200: the image branch runs on CUDA only (no CPU fallback)")
        return self.gather_rows(rgb, choose).transpose(1, 2).contiguous()

    def gather_rows(self, rgb, choose):
        """Same as gather() but channels-last: (B,N,128), the layout the per-point MLP kernels consume."""
        if not rgb.is_cuda:
            raise RuntimeError("istnet_b200: the image branch runs on CUDA only (no CPU fallback)")
        from .image_engine import image_branch

        return image_branch(self.model, rgb, choose)
