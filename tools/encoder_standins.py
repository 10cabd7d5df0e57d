"""Torch stand-ins for the encoders the polarization path feeds -- MEASUREMENT ONLY (BASELINE configs[3], SURVEY 8d:
"time kernel alone and kernel + encoders").  They reproduce the layer geometry of the reference's modules so that
their forward time on B200 can be put next to the polcue kernel's; weights are random, nothing is trained, and no
arithmetic of the polarization path lives here.

  XOLP branch     ShallowEncoder('XOLP')      manydepth/networks/pre_encoders.py:49-83   2 -> 64 channels, /8
  normals branch  ShallowNormalsEncoder       manydepth/networks/pre_encoders.py:85-97   9 -> 64 channels, /8 (get_normals is polcue's)
  RGB branch      ShallowResnetEncoder(18)    manydepth/networks/resnet_encoder.py:783-822  resnet18 stem + layer1 + layer2
"""
import torch
import torch.nn as nn


def _unit(cin, cout, k, down):
    """conv (stride 2 when down == 'stride') -> batch norm -> ReLU -> optional 2x2 max pool (dropout is identity in eval mode)."""
    layers = [nn.Conv2d(cin, cout, k, stride=2 if down == "stride" else 1, padding=k // 2), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
    if down == "pool":
        layers.append(nn.MaxPool2d(2))
    return nn.Sequential(*layers)


class _Residual(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.body = nn.Sequential(_unit(ch, ch, 3, None), _unit(ch, ch, 3, None))

    def forward(self, x):
        return self.body(x) + x


def shallow_branch(cin):
    """7x7 stride-2 unit, then twice (5x5 unit + max pool), a two-conv residual block after each: cin x H x W -> 64 x H/8 x W/8."""
    return nn.Sequential(_unit(cin, 64, 7, "stride"), _Residual(64), _unit(64, 64, 5, "pool"), _Residual(64), _unit(64, 64, 5, "pool"),
                         _Residual(64))


class PolarEncoders(nn.Module):
    XOLP_MEAN, XOLP_STD = 0.08693199701957657, 0.44430732785457433      # pre_encoders.py:79

    def __init__(self):
        super().__init__()
        import torchvision
        self.xolp_branch = shallow_branch(2)
        self.normals_branch = shallow_branch(9)
        r18 = torchvision.models.resnet18(weights=None)
        self.rgb_stem = nn.Sequential(r18.conv1, r18.bn1, r18.relu)
        self.rgb_tail = nn.Sequential(r18.maxpool, r18.layer1, r18.layer2)

    def forward(self, xolp, normals, rgb):
        fx = self.xolp_branch((xolp - self.XOLP_MEAN) / self.XOLP_STD)
        fn = self.normals_branch(normals)
        fr = self.rgb_tail(self.rgb_stem((rgb - 0.45) / 0.225))
        return fx, fn, fr
