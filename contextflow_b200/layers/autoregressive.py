"""Masked (autoregressive) convolution blocks of `--coupling maf` (reference layers/autoregressive/masked_conv_2d.py:7-99,
layers/autoregressive/utils.py:25-92): the same module tree and buffer names (`conv{1,2,3}.weight/bias/mask`), so reference checkpoints
load; the arithmetic runs in libcfpp (three conv launches with ReLU on the input over mask-multiplied weights)."""
import torch
import torch.nn as nn

from .. import ops

__all__ = ['MaskedConv2d', 'MaskedResidualBlock2d', 'MaskedLinear', 'MaskedResidualBlockLinear', 'mask_channels', 'mask_conv2d']


def mask_channels(mask_type, in_channels, out_channels, data_channels=3):
    """utils.py:25-57: the (data_channels x data_channels) lower-triangular base ('A': strict) tiled over the feature maps."""
    base = torch.ones(data_channels, data_channels).tril(-1 if mask_type == 'A' else 0)
    rows = torch.cat([base] * (in_channels // data_channels + 1), dim=1)
    full = torch.cat([rows] * (out_channels // data_channels + 1), dim=0)
    return full[:out_channels, :in_channels]


def mask_conv2d(mask_type, in_channels, out_channels, height, width, data_channels=3):
    """utils.py:60-92: channel mask at the central tap, nothing to the right of or below the centre."""
    mask = torch.ones(out_channels, in_channels, height, width)
    mask[:, :, height // 2, width // 2] = mask_channels(mask_type, in_channels, out_channels, data_channels)
    mask[:, :, height // 2, width // 2 + 1:] = 0
    mask[:, :, height // 2 + 1:] = 0
    return mask


class MaskedConv2d(nn.Conv2d):
    def __init__(self, *args, mask_type, data_channels=3, **kwargs):
        super().__init__(*args, **kwargs)
        assert mask_type in {'A', 'B'}
        o, i, h, w = self.weight.size()
        self.register_buffer('mask', mask_conv2d(mask_type, i, o, h, w, data_channels))
        self._masked_version = None

    def masked_weight(self):
        """masked_conv_2d.py:21-23 multiplies `weight.data` by the mask in place on every forward; once per weight version is the
        same state (the product is idempotent)."""
        from .flowlayer import derived_is_live
        key = (self.weight.data_ptr(), self.weight._version)
        if self._masked_version != key or derived_is_live([self.weight]):
            with torch.no_grad():
                self.weight.data.mul_(self.mask)
            self._masked_version = (self.weight.data_ptr(), self.weight._version)
        return self.weight.detach()

    def forward(self, x):
        return ops.conv2d_fwd(x, x.shape[1], self.masked_weight(), self.bias.detach(), relu=False)


class MaskedResidualBlock2d(nn.Module):
    """masked_conv_2d.py:81-98: conv1(relu(x)) -> conv2(relu(.)) -> conv3(relu(.)) + cat(x, x).  `forward(x, identity=False)` returns the
    conv output without the identity: the fused coupling kernel adds x itself."""

    def __init__(self, I, O, kernel_size=(1, 1), padding=(0, 0), D=0, mask_type='B'):
        super().__init__()
        self.conv1 = MaskedConv2d(1 * I, 2 * I, 1, mask_type=mask_type, data_channels=D)
        self.conv2 = MaskedConv2d(2 * I, 2 * I, kernel_size, padding=padding, padding_mode='reflect', mask_type=mask_type, data_channels=D)
        self.conv3 = MaskedConv2d(2 * I, 2 * O, 1, mask_type=mask_type, data_channels=D)
        ks, pd = tuple(self.conv2.kernel_size), tuple(self.conv2.padding)
        if pd != (ks[0] // 2, ks[1] // 2) or any(k not in (1, 3) for k in ks):
            raise NotImplementedError('the conv kernels assume "same" reflect padding with 1/3-wide kernels (model.py:114)')
        from .flowlayer import PackCache
        self._packs = PackCache()

    def _tc_pack(self):
        """The block is the conditioner's 1x1 -> ReLU -> kxk reflect -> ReLU -> 1x1 chain applied to relu(x): same tcgen05 kernel, weights
        packed from the mask-multiplied tensors (once per weight version)."""
        convs = (self.conv1, self.conv2, self.conv3)
        ws = [c.masked_weight() for c in convs]
        if not ws[0].is_cuda:
            return None
        key = [c.weight for c in convs] + [c.bias for c in convs]
        return self._packs.get('tc', key, lambda: (ops.conv_cond_tc_pack(ws[0], ws[1], ws[2], ws[0].shape[1]),
                                                   ops.pad_vec(self.conv1.bias), ops.pad_vec(self.conv2.bias), ops.pad_vec(self.conv3.bias)))

    def forward(self, x, identity=True):
        if identity:
            raise NotImplementedError('use MaskedCoupling: the identity is added inside the fused coupling kernel')
        pk = self._tc_pack()
        if pk is not None and pk[0] is not None and ops.conv_cond_tc_mode() != 'fma':
            ks = self.conv2.kernel_size
            h = ops.conv_cond_tc(ops.relu(x), x.shape[1], pk[0], pk[1], pk[2], pk[3], self.conv2.out_channels, x.shape[2], x.shape[3],
                                 ks[0], ks[1], self.conv3.out_channels)
            if h is not None:
                return h
        h = x
        for conv in (self.conv1, self.conv2, self.conv3):                  # FP32 kernels: shapes without a tensor-core plan
            h = ops.conv2d_fwd(h, h.shape[1], conv.masked_weight(), conv.bias.detach(), relu=False, relu_in=True)
        return h


class MaskedLinear(nn.Linear):
    """masked_linear.py:17-102, the hidden-layer form MaskedResidualBlockLinear uses (is_output = False, no random mask): out degrees
    cycle over 1 .. data_features - 1, mask[o, i] = out_degree[o] >= in_degree[i]; forward = F.linear(x, weight * mask, bias).  Buffers
    `mask` and `degrees` as in the reference, so checkpoints load."""

    def __init__(self, in_degrees, out_features, data_features, bias=True):
        super().__init__(in_features=len(in_degrees), out_features=out_features, bias=bias)
        self.data_features = data_features
        max_, min_ = max(1, data_features - 1), min(1, data_features - 1)
        out_degrees = torch.arange(out_features) % max_ + min_
        self.register_buffer('mask', (out_degrees[..., None] >= in_degrees).float())
        self.register_buffer('degrees', out_degrees)

    @staticmethod
    def get_data_degrees(in_features):
        return torch.arange(1, in_features + 1)

    def masked_weight(self):
        return (self.weight * self.mask).detach()

    def forward(self, x):
        return ops.rows_linear(x, self.masked_weight(), self.bias.detach())


class MaskedResidualBlockLinear(nn.Module):
    """masked_linear.py:104-128: linear1(relu(x)) -> linear2(relu(.)) -> linear3(relu(.)) + x.  The output is 2*O wide and the identity I
    wide: like the reference's `x + identity`, this only works when I == 2*O or I == 1."""

    def __init__(self, I, O, D):
        super().__init__()
        self.linear1 = MaskedLinear(MaskedLinear.get_data_degrees(1 * I), 2 * I, D)
        self.linear2 = MaskedLinear(MaskedLinear.get_data_degrees(2 * I), 2 * I, D)
        self.linear3 = MaskedLinear(MaskedLinear.get_data_degrees(2 * I), 2 * O, D)
        self.I, self.O = I, O

    def forward(self, c):
        if self.I not in (1, 2 * self.O):
            raise RuntimeError(f'The size of tensor a ({2 * self.O}) must match the size of tensor b ({self.I}) at non-singleton dimension 1')   # masked_linear.py:128
        h = c
        for lin in (self.linear1, self.linear2, self.linear3):
            h = lin(ops.relu(h))
        ident = c if self.I == 2 * self.O else c.expand(c.shape[0], 2 * self.O).contiguous()
        return ops.add(h, ident)
