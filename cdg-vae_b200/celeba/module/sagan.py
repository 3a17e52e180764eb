"""Parameter containers mirroring the reference's celeba/module/sagan.py Generator (SAGAN generator, 128 px):
same module tree, hence the same `state_dict()` keys (`block1.conv_1.weight_orig`, `...weight_u`, `bn.running_var`,
...), and the same construction order, hence identical tensors under the same seed.  The arithmetic of these
layers runs in libcdgvae_sm100.so (csrc/celeba_step.cu); the modules here only own the tensors.

    Generator(latent_dim, conv_dim=32, image_size=128, out_channels=3, add_noise=True, attn=True)   sagan.py:137-210
"""
import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm


class _Holder(nn.Module):
    """A module that only holds parameters / sub-modules; evaluation happens in the CUDA library."""

    def forward(self, *a, **k):
        raise RuntimeError("this layer is evaluated inside libcdgvae_sm100.so: call CDGVAE.decode()/forward()")


class NoiseInjection(_Holder):                                  # sagan.py:74-84 (weight stays 0: an exact no-op)
    def __init__(self, channel, size=1):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1, channel, size, size))


def _sn_conv(cin, cout, k, pad):
    return spectral_norm(nn.Conv2d(cin, cout, k, 1, pad))


class GenIniBlock(_Holder):                                     # sagan.py:86-101
    def __init__(self, z_dim, out_channels, size=1, add_noise=True):
        super().__init__()
        self.out_channels = out_channels
        self.snlinear0 = spectral_norm(nn.Linear(z_dim, out_channels * 16))
        if add_noise:
            self.noise0 = NoiseInjection(out_channels, size)


class GenBlock(_Holder):                                        # sagan.py:103-140
    def __init__(self, in_channels, out_channels, size=1, add_noise=True):
        super().__init__()
        self.conv_1 = _sn_conv(in_channels, out_channels, 3, 1)
        self.conv_2 = _sn_conv(out_channels, out_channels, 3, 1)
        if add_noise:
            self.noise1 = NoiseInjection(out_channels, size)
            self.noise2 = NoiseInjection(out_channels, size)
        self.conv_0 = _sn_conv(in_channels, out_channels, 1, 0)
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)


class Self_Attn(_Holder):                                       # sagan.py:31-73 (sigma stays 0: an exact no-op)
    def __init__(self, in_channels):
        super().__init__()
        c = in_channels
        self.snconv1x1_theta = _sn_conv(c, c // 8, 1, 0)
        self.snconv1x1_phi = _sn_conv(c, c // 8, 1, 0)
        self.snconv1x1_g = _sn_conv(c, c // 2, 1, 0)
        self.snconv1x1_attn = _sn_conv(c // 2, c, 1, 0)
        self.sigma = nn.Parameter(torch.zeros(1))


class Generator(_Holder):
    def __init__(self, latent_dim, conv_dim=32, image_size=128, out_channels=3, add_noise=True, attn=True):
        super().__init__()
        if image_size != 128 or conv_dim != 32 or out_channels != 3 or not add_noise or not attn:
            raise ValueError("only the configuration celeba/module/model.py uses is built: Generator(z, 32, 128, 3, True, True)")
        self.latent_dim, self.conv_dim, self.image_size = latent_dim, conv_dim, image_size
        c = conv_dim
        self.block0 = GenIniBlock(latent_dim, c * 16, 4)
        self.block1 = GenBlock(c * 16, c * 16, size=8)
        self.block2 = GenBlock(c * 16, c * 8, size=16)
        self.block3 = GenBlock(c * 8, c * 4)
        self.self_attn1 = Self_Attn(c * 4)
        self.block4 = GenBlock(c * 4, c * 2)
        self.block5 = GenBlock(c * 2, c)
        self.bn = nn.BatchNorm2d(c, eps=1e-5, momentum=0.0001, affine=True)
        self.toRGB = _sn_conv(c, out_channels, 3, 1)
        # sagan.py:190 `self.apply(init_weights)`: orthogonal_ lands on the derived `.weight` attribute of the
        # spectral-norm wrapped layers (overwritten by the next training-mode forward) but advances the RNG; biases -> 0
        for m in self.modules():
            if type(m) in (nn.Linear, nn.Conv2d):
                nn.init.orthogonal_(m.weight)
                m.bias.data.fill_(0.0)

    def sn_layers(self):
        """Spectral-norm layers in evaluation order: lin0, (conv_1, conv_2, conv_0) x 5, 4 attention convs, toRGB."""
        out = [("block0.snlinear0", self.block0.snlinear0)]
        for b in range(1, 6):
            blk = getattr(self, f"block{b}")
            out += [(f"block{b}.conv_1", blk.conv_1), (f"block{b}.conv_2", blk.conv_2), (f"block{b}.conv_0", blk.conv_0)]
        out += [(f"self_attn1.snconv1x1_{n}", getattr(self.self_attn1, f"snconv1x1_{n}")) for n in ("theta", "phi", "g", "attn")]
        out.append(("toRGB", self.toRGB))
        return out
