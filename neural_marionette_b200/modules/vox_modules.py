"""3-D conv building blocks of the detector (drop-in for the reference's
modules/vox_modules.py: same class names, constructor arguments and parameter
names/shapes, so reference checkpoints load with strict=True).

Each class is a parameter container (torch.nn layers own the weights) whose
arithmetic runs through the nm_b200 kernels on fp16 channels-last activations:
``run(x_act)`` is what the detector calls; ``forward(x)`` keeps the reference's
NCDHW fp32 tensor contract for callers that use a block on its own.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import autograd as AG
from .. import ops


def _norm(channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(channels // 16, channels)


class _ActModule(nn.Module):
    """forward() = NCDHW fp32 in/out around run() (no autograd graph: the training step goes through
    KyptDetector.forward, which calls run_train())."""

    def forward(self, x):
        with torch.no_grad():
            return ops.act_to_ncdhw(self.run(ops.ncdhw_to_act(x)))


def conv_gn_lrelu(x, conv, gn, in_affine=None):
    """LeakyReLU(GN(conv(input))).  Without in_affine the input is x.  With in_affine = (scale, shift, act[, x2, scale2,
    shift2]) x is a RAW conv output and input = act(x*scale+shift) [+ x2*scale2+shift2]; that is applied inside this
    conv's operand path when the kernel supports it, else by a separate pass."""
    if in_affine is not None:
        dual = len(in_affine) > 3 and in_affine[3] is not None
        if ops.can_fuse_input2(x, conv) if dual else ops.can_fuse_input(x, conv):
            raw, a, b = ops.conv3d(x, conv, gn, in_affine=in_affine)
            return ops.affine_act(raw, a, b, True)
        if dual:
            x = ops.affine_act(x, in_affine[0], in_affine[1], in_affine[2], x2=in_affine[3], a2=in_affine[4],
                               b2=in_affine[5])
        else:
            x = ops.affine_act(x, in_affine[0], in_affine[1], in_affine[2])
    raw, a, b = ops.conv3d(x, conv, gn)
    return ops.affine_act(raw, a, b, True)


class Basic3DBlock(_ActModule):
    """conv(k, pad (k-1)//2) -> GroupNorm(C//16) -> LeakyReLU  (reference vox_modules.py:8-19)."""

    def __init__(self, in_planes, out_planes, kernel_size):
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=1, padding=(kernel_size - 1) // 2),
            _norm(out_planes), nn.LeakyReLU())

    def run(self, x):
        return conv_gn_lrelu(x, self.block[0], self.block[1])

    def run_train(self, x):
        return AG.conv_gn_act(x, self.block[0], self.block[1], True)

    def run_coordconv_train(self, occ):
        return AG.first_conv_gn_act(occ, self.block[0], self.block[1])

    def run_coordconv(self, occ):
        """occ (n, G, G, G) fp32 occupancy; the 3 coordinate channels are synthesised in-kernel."""
        raw, a, b = self.run_coordconv_raw(occ)
        return ops.affine_act(raw, a, b, True)

    def run_coordconv_raw(self, occ):
        """-> (raw conv output, GroupNorm scale, shift): the consumer applies the normalisation + LeakyReLU."""
        return ops.first_conv(occ, self.block[0], self.block[1])


class Res3DBlock(_ActModule):
    """[conv3,GN,LReLU,conv3,GN](x) + skip(x); skip = identity or conv1+GN.  The reference's trailing
    ``F.leaky_relu(., True)`` has slope 1.0, i.e. no activation (reference vox_modules.py:22-47)."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.res_branch = nn.Sequential(
            nn.Conv3d(in_planes, out_planes, kernel_size=3, stride=1, padding=1), _norm(out_planes), nn.LeakyReLU(),
            nn.Conv3d(out_planes, out_planes, kernel_size=3, stride=1, padding=1), _norm(out_planes))
        if in_planes == out_planes:
            self.skip_con = nn.Sequential()
        else:
            self.skip_con = nn.Sequential(
                nn.Conv3d(in_planes, out_planes, kernel_size=1, stride=1, padding=0), _norm(out_planes))

    def run(self, x):
        raw, a, b, x2, a2, b2 = self.run_raw(x)
        return ops.affine_act(raw, a, b, False, x2=x2, a2=a2, b2=b2)

    def run_train(self, x):
        """Training-mode forward (autograd Functions; activations are materialised for the backward)."""
        h = AG.conv_gn_act(x, self.res_branch[0], self.res_branch[1], True)
        skip = x if len(self.skip_con) == 0 else AG.conv_gn_act(x, self.skip_con[0], self.skip_con[1], False)
        return AG.conv_gn_act(h, self.res_branch[3], self.res_branch[4], False, res=skip)

    def run_raw(self, x):
        """-> (raw, scale, shift, x2, scale2, shift2) with block output = raw*scale+shift + x2*scale2+shift2 (x2 as is
        when scale2 is None): a consumer that can apply this in its operand path avoids one pass over the tensor."""
        raw1, a1, b1 = ops.conv3d(x, self.res_branch[0], self.res_branch[1])
        if ops.can_fuse_input(raw1, self.res_branch[3]):
            # GroupNorm + LeakyReLU of the first conv applied inside the second conv's operand path
            raw, a, b = ops.conv3d(raw1, self.res_branch[3], self.res_branch[4], in_affine=(a1, b1, True))
        else:
            raw, a, b = ops.conv3d(ops.affine_act(raw1, a1, b1, True), self.res_branch[3], self.res_branch[4])
        if len(self.skip_con) == 0:
            return raw, a, b, x, None, None
        sraw, sa, sb = ops.conv3d(x, self.skip_con[0], self.skip_con[1])
        return raw, a, b, sraw, sa, sb


class Pool3DBlock(_ActModule):
    """Learned 2x down-sample: conv(k2, s2) -> GN -> LReLU (reference vox_modules.py:49-61)."""

    def __init__(self, pool_size, input_plane):
        super().__init__()
        self.stride_conv = nn.Sequential(
            nn.Conv3d(input_plane, input_plane, kernel_size=pool_size, stride=pool_size, padding=0),
            _norm(input_plane), nn.LeakyReLU())

    def run(self, x, in_affine=None):
        return conv_gn_lrelu(x, self.stride_conv[0], self.stride_conv[1], in_affine)

    def run_train(self, x):
        return AG.conv_gn_act(x, self.stride_conv[0], self.stride_conv[1], True)


class Upsample3DBlock(_ActModule):
    """ConvTranspose3d(k2, s2) -> GN -> LReLU (reference vox_modules.py:63-75)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride, output_padding=0):
        super().__init__()
        assert stride == 2
        self.block = nn.Sequential(
            nn.ConvTranspose3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=0,
                               output_padding=output_padding),
            _norm(out_planes), nn.LeakyReLU())

    def run(self, x, skip=None):
        raw, a, b = ops.conv_transpose3d(x, self.block[0], self.block[1])
        return ops.affine_act(raw, a, b, True, x2=skip)

    def run_train(self, x, skip=None):
        return AG.conv_gn_act(x, self.block[0], self.block[1], True, res=skip)


class HG(_ActModule):
    """3-level hour-glass, channels in->32->48->72->72->48->32->out, additive skips
    (reference vox_modules.py:78-120)."""

    def __init__(self, input_channels, output_channels, N=88):
        super().__init__()
        outer_padding = [(N // 4) % 2, (N // 2) % 2, N % 2]
        self.encoder_pool1 = Pool3DBlock(2, input_channels)
        self.encoder_res1 = Res3DBlock(input_channels, 32)
        self.encoder_pool2 = Pool3DBlock(2, 32)
        self.encoder_res2 = Res3DBlock(32, 48)
        self.encoder_pool3 = Pool3DBlock(2, 48)
        self.encoder_res3 = Res3DBlock(48, 72)
        self.decoder_res3 = Res3DBlock(72, 72)
        self.decoder_upsample3 = Upsample3DBlock(72, 48, 2, 2, outer_padding[0])
        self.decoder_res2 = Res3DBlock(48, 48)
        self.decoder_upsample2 = Upsample3DBlock(48, 32, 2, 2, outer_padding[1])
        self.decoder_res1 = Res3DBlock(32, 32)
        self.decoder_upsample1 = Upsample3DBlock(32, output_channels, 2, 2, outer_padding[2])
        self.skip_res1 = Res3DBlock(input_channels, output_channels)
        self.skip_res2 = Res3DBlock(32, 32)
        self.skip_res3 = Res3DBlock(48, 48)

    def run(self, x):
        s1 = self.skip_res1.run(x)
        x = self.encoder_res1.run(self.encoder_pool1.run(x))
        s2 = self.skip_res2.run(x)
        x = self.encoder_res2.run(self.encoder_pool2.run(x))
        s3 = self.skip_res3.run(x)
        x = self.encoder_res3.run(self.encoder_pool3.run(x))
        x = self.decoder_res3.run(x)
        x = self.decoder_res2.run(self.decoder_upsample3.run(x, skip=s3))
        x = self.decoder_res1.run(self.decoder_upsample2.run(x, skip=s2))
        return self.decoder_upsample1.run(x, skip=s1)

    def run_train(self, x):
        s1 = self.skip_res1.run_train(x)
        x = self.encoder_res1.run_train(self.encoder_pool1.run_train(x))
        s2 = self.skip_res2.run_train(x)
        x = self.encoder_res2.run_train(self.encoder_pool2.run_train(x))
        s3 = self.skip_res3.run_train(x)
        x = self.encoder_res3.run_train(self.encoder_pool3.run_train(x))
        x = self.decoder_res3.run_train(x)
        x = self.decoder_res2.run_train(self.decoder_upsample3.run_train(x, skip=s3))
        x = self.decoder_res1.run_train(self.decoder_upsample2.run_train(x, skip=s2))
        return self.decoder_upsample1.run_train(x, skip=s1)
