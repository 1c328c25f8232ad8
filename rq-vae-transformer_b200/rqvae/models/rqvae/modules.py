"""Parameter trees of the conv encoder / decoder.

These modules only HOLD parameters under the reference's state_dict keys (reference builders:
rqvae/models/rqvae/modules.py:10-70,101-168 and layers.py:16-157) -- they have no ``forward``: the layer plan is
executed by the native engine (csrc/vae_engine.cu), which resolves the same key names.  Construction order follows
the reference so that seeded default initialisation produces the same weights."""
from torch import nn


def _norm(ch):
    return nn.GroupNorm(num_groups=32, num_channels=ch, eps=1e-6, affine=True)


def _conv(cin, cout, k, stride=1, pad=0):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad)


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("rqb200: parameter holder -- compute runs in the native engine (RQVAE.encode/decode)")


class ResnetBlock(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.checkpointing = False
        self.norm1 = _norm(cin)
        self.conv1 = _conv(cin, cout, 3, 1, 1)
        self.norm2 = _norm(cout)
        self.conv2 = _conv(cout, cout, 3, 1, 1)
        if cin != cout:
            self.nin_shortcut = _conv(cin, cout, 1)


class AttnBlock(_Holder):
    def __init__(self, ch):
        super().__init__()
        self.in_channels = ch
        self.norm = _norm(ch)
        self.q, self.k, self.v = _conv(ch, ch, 1), _conv(ch, ch, 1), _conv(ch, ch, 1)
        self.proj_out = _conv(ch, ch, 1)


class _Resample(_Holder):
    def __init__(self, ch, stride):
        super().__init__()
        self.with_conv = True
        self.conv = _conv(ch, ch, 3, stride, 1 if stride == 1 else 0)


class _Level(_Holder):
    pass


def _mid(ch):
    mid = _Level()
    mid.block_1 = ResnetBlock(ch, ch)
    mid.attn_1 = AttnBlock(ch)
    mid.block_2 = ResnetBlock(ch, ch)
    return mid


class Encoder(_Holder):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, **ignored):
        super().__init__()
        if not resamp_with_conv:
            raise NotImplementedError("rqb200: avg-pool downsampling is not implemented (no shipped config uses it)")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = _conv(in_channels, ch, 3, 1, 1)
        res, width = resolution, ch
        self.down = nn.ModuleList()
        for lvl, mult in enumerate(ch_mult):
            level = _Level()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks):
                level.block.append(ResnetBlock(width, ch * mult))
                width = ch * mult
                if res in attn_resolutions:
                    level.attn.append(AttnBlock(width))
            if lvl != len(ch_mult) - 1:
                level.downsample = _Resample(width, 2)
                res //= 2
            self.down.append(level)
        self.mid = _mid(width)
        self.norm_out = _norm(width)
        self.conv_out = _conv(width, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class Decoder(_Holder):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, **ignored):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        width = ch * ch_mult[-1]
        res = resolution // 2 ** (len(ch_mult) - 1)
        self.z_shape = (1, z_channels, res, res)
        self.conv_in = _conv(z_channels, width, 3, 1, 1)
        self.mid = _mid(width)
        levels = []
        for lvl in reversed(range(len(ch_mult))):
            level = _Level()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                level.block.append(ResnetBlock(width, ch * ch_mult[lvl]))
                width = ch * ch_mult[lvl]
                if res in attn_resolutions:
                    level.attn.append(AttnBlock(width))
            if lvl != 0:
                level.upsample = _Resample(width, 1)
                res *= 2
            levels.insert(0, level)
        self.up = nn.ModuleList(levels)
        self.norm_out = _norm(width)
        self.conv_out = _conv(width, out_ch, 3, 1, 1)
