"""UNet on a ResNet-34 encoder: the 2D network in front of the hot path (reference:
mvpnet/models/unet_resnet34.py:9-125).  Dense cuDNN convolutions, out of scope for hand-written
kernels but inside the measured MVPNet forward, so it is provided here with the same
constructor, forward contract ({'image'} -> {'seg_logit', 'feature'}) and state_dict keys.

The first convolution has stride 1 (weights shaped like torchvision's conv1); inputs are zero-padded
to a multiple of 16 and the output is cropped back.  `feature` is the 64-channel decoder output at
input resolution.
"""
import torch
from torch import nn
import torch.nn.functional as F
from torchvision.models.resnet import resnet34


def _up(c_in, c_out):
    return nn.Sequential(nn.ConvTranspose2d(c_in, c_out, kernel_size=2, stride=2), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True))


def _fuse(c_in, c_out):
    return nn.Sequential(nn.Conv2d(c_in, c_out, kernel_size=3, padding=1), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True))


class UNetResNet34(nn.Module):
    def __init__(self, num_classes, p=0.0, pretrained=True):
        super().__init__()
        self.num_classes = num_classes
        net = resnet34(weights='IMAGENET1K_V1' if pretrained else None)
        self.encoder0 = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3, bias=False)
        self.encoder0.weight.data = net.conv1.weight.data
        self.bn, self.relu, self.maxpool = net.bn1, net.relu, net.maxpool
        self.encoder1, self.encoder2, self.encoder3, self.encoder4 = net.layer1, net.layer2, net.layer3, net.layer4
        self.deconv4, self.decoder3 = _up(512, 256), _fuse(512, 256)
        self.deconv3, self.decoder2 = _up(256, 128), _fuse(256, 128)
        self.deconv2, self.decoder1 = _up(128, 64), _fuse(128, 64)
        self.deconv1, self.decoder0 = _up(64, 64), _fuse(128, 64)
        self.logit = nn.Conv2d(64, num_classes, 1, bias=True)
        self.dropout = nn.Dropout(p=p) if p > 0.0 else None

    def features(self, x):
        """image (n,3,h,w) -> 64-channel feature map (n,64,h,w)."""
        h, w = x.shape[2], x.shape[3]
        pad_h, pad_w = (-h) % 16, (-w) % 16
        if pad_h or pad_w:
            x = F.pad(x, [0, pad_w, 0, pad_h])
        skips = []
        x = self.relu(self.bn(self.encoder0(x)))
        skips.append(x)
        x = self.encoder1(self.maxpool(x))
        skips.append(x)
        x = self.encoder2(x)
        skips.append(x)
        x = self.encoder3(x)
        if self.dropout is not None:
            x = self.dropout(x)
        skips.append(x)
        x = self.encoder4(x)
        if self.dropout is not None:
            x = self.dropout(x)
        for up, fuse, skip in ((self.deconv4, self.decoder3, skips[3]), (self.deconv3, self.decoder2, skips[2]),
                               (self.deconv2, self.decoder1, skips[1]), (self.deconv1, self.decoder0, skips[0])):
            x = fuse(torch.cat([up(x), skip], dim=1))
        if pad_h or pad_w:
            x = x[:, :, 0:h, 0:w]
        return x

    def forward(self, data_dict):
        x = self.features(data_dict['image'])
        return {'seg_logit': self.logit(x), 'feature': x}
