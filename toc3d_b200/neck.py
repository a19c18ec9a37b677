"""Drop-in `img_neck` plugin: CPFPN on the sm_100a GEMM (the first "next" row).

Reference: projects/mmdet3d_plugin/models/necks/cp_fpn.py:11-208, shipped config
`img_neck=dict(type='CPFPN', in_channels=[1024], out_channels=256, num_outs=2)` (ToC3D_fast.py:70-74):
lateral 1x1 conv 1024 -> 256, 3x3 conv 256 -> 256 (pad 1) on level 0, second level =
`max_pool2d(kernel=1, stride=2)` of the first (a strided subsample).  Same registry name (`CPFPN` in
mmdet's NECKS), constructor kwargs, forward(list of NCHW maps) -> tuple of NCHW fp32 maps, and
state-dict keys (`lateral_convs.0.conv.*`, `fpn_convs.0.conv.*`).  Only the shipped single-level
configuration is implemented; anything else raises.  Inference only, CUDA only, no fallback.
"""
import torch
import torch.nn as nn

from . import lib as L

try:
    from mmdet.models import NECKS as _NECKS

    def _register(cls):
        return _NECKS.register_module(force=True)(cls)
except Exception:
    def _register(cls):
        return cls


class _ConvModule(nn.Module):
    """mmcv ConvModule with norm_cfg=None, act_cfg=None: a Conv2d with bias under the name `conv`."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=True)
        nn.init.xavier_uniform_(self.conv.weight)          # init_cfg=dict(type='Xavier', layer='Conv2d', distribution='uniform')
        nn.init.zeros_(self.conv.bias)


@_register
class CPFPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode="nearest"), init_cfg=None):
        super().__init__()
        assert isinstance(in_channels, list)
        if (len(in_channels) != 1 or start_level != 0 or end_level not in (-1, 1) or add_extra_convs or conv_cfg or norm_cfg
                or act_cfg or num_outs < 1):
            raise NotImplementedError("only the shipped CPFPN configuration (one input level, no norm / activation / "
                                      "extra convs) is implemented")
        if in_channels[0] % 64 or out_channels % 64:
            raise NotImplementedError("channel counts must fit the GEMM k-blocks (in %% 64 == 0, out %% 64 == 0)")
        self.in_channels, self.out_channels, self.num_outs = in_channels, out_channels, num_outs
        self.fp16_enabled = False
        self.lateral_convs = nn.ModuleList([_ConvModule(in_channels[0], out_channels, 1)])
        self.fpn_convs = nn.ModuleList([_ConvModule(out_channels, out_channels, 3, padding=1)])
        self._packed = None
        self._layouts = {}

    def _invalidate(self):
        self._packed = None
        for ref in getattr(self, "_fused_into", []):       # backbones whose captured graphs hold pointers to the packed weights
            bb = ref()
            if bb is not None:
                bb._graphs = {}

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def _weights(self, dev):
        if self._packed is None or self._packed["dev"] != dev:
            co = self.out_channels
            w1 = self.lateral_convs[0].conv.weight.detach().to(dev).float().reshape(co, -1)
            # [co, ci, ky, kx] -> [co, (ky, kx, ci)]: tap-major weight columns of the implicit 3x3 GEMM
            w3 = self.fpn_convs[0].conv.weight.detach().to(dev).float().permute(0, 2, 3, 1).reshape(co, -1)
            self._packed = dict(dev=dev, w1=w1.to(torch.bfloat16).contiguous(), w3=w3.to(torch.bfloat16).contiguous(),
                                b1=self.lateral_convs[0].conv.bias.detach().to(dev).float().contiguous(),
                                b3=self.fpn_convs[0].conv.bias.detach().to(dev).float().contiguous())
        return self._packed

    def _conv_layout(self, dev, V, H, W):
        """Spatially zero-padded NHWC layout of the lateral map ((H+2) x (W+2) pixels per image) for the implicit 3x3
        convolution: row maps between the V*H*W pixel rows and the padded rows, the tap row shifts, and the padded
        bf16 buffer itself (borders are zero and never written).  Cached per (device, V, H, W)."""
        key = (str(dev), V, H, W)
        c = self._layouts.get(key)
        if c is None:
            Hp, Wp = H + 2, W + 2
            v, y, x = torch.meshgrid(torch.arange(V), torch.arange(H), torch.arange(W), indexing="ij")
            to_pad = ((v * Hp + y + 1) * Wp + x + 1).reshape(-1).to(torch.int32)
            from_pad = torch.full((V * Hp * Wp,), -1, dtype=torch.int32)
            from_pad[to_pad.long()] = torch.arange(V * H * W, dtype=torch.int32)
            c = dict(Mp=V * Hp * Wp, to_pad=to_pad.to(dev), from_pad=from_pad.to(dev),
                     shifts=[(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)],
                     lat=torch.zeros(V * Hp * Wp, self.out_channels, device=dev, dtype=torch.bfloat16))
            self._layouts[key] = c
        return c

    def launch(self, x2d, V, H, W):
        """The neck's launch sequence on the residual-stream buffer itself: x2d = fp32 (or bf16) [V*H*W, C] NHWC rows.
        -> level-0 output, fp32 [V*H*W, out_channels] (NHWC rows).  Stream-ordered, allocation-only host work, so the
        backbone can run it inside its own CUDA graph (ToC3DEVAViT.fuse_neck).
        Lateral 1x1 conv (cp_fpn.py:163-166) = GEMM whose epilogue scatters the rows into the zero-padded layout; 3x3 conv
        (cp_fpn.py:182-184) = IMPLICIT GEMM over that layout: the nine taps are nine row-shifted TMA views of the same
        matrix (toc3d_epilogue.conv_*), no im2col buffer; border rows of the result are dropped by the output map."""
        M, C = x2d.shape
        co = self.out_channels
        p = self._weights(x2d.device)
        lay = self._conv_layout(x2d.device, V, H, W)
        if x2d.dtype == torch.bfloat16:
            a = x2d
        else:
            a = torch.empty(M, C, device=x2d.device, dtype=torch.bfloat16)
            L.cast_bf16(x2d, a)
        L.gemm(a, p["w1"], L.EPI_LINEAR, bias=p["b1"], out=lay["lat"], out_map=lay["to_pad"])
        out0 = torch.empty(M, co, device=x2d.device, dtype=torch.float32)
        L.gemm(lay["lat"], p["w3"], L.EPI_LINEAR, M=lay["Mp"], bias=p["b3"], out=out0, out_f32=True, out_map=lay["from_pad"],
               conv_cin=co, conv_row_shift=lay["shifts"])
        return out0

    def levels(self, out0, V, H, W):
        """NCHW views of the output pyramid: level 0 and its stride-2 subsamples (F.max_pool2d(x, 1, stride=2),
        cp_fpn.py:190-191)."""
        outs = [out0.view(V, H, W, self.out_channels).permute(0, 3, 1, 2)]
        for _ in range(self.num_outs - 1):
            outs.append(outs[-1][:, :, ::2, ::2])
        return tuple(outs)

    @torch.no_grad()
    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)                                       # cp_fpn.py:160
        x = inputs[0]
        fused = getattr(x, "_toc3d_fused_neck", None)
        if fused is not None and fused[0] is self:
            return fused[1]                 # computed inside the backbone's CUDA graph (ToC3DEVAViT.fuse_neck)
        if not x.is_cuda:
            raise RuntimeError("toc3d_b200 CPFPN runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training:
            raise RuntimeError("toc3d_b200 CPFPN is inference-only; call .eval()")
        V, C, H, W = x.shape
        nhwc = x.permute(0, 2, 3, 1)                        # the backbone hands out a permuted view of NHWC storage
        nhwc = nhwc if nhwc.is_contiguous() else nhwc.contiguous()
        if nhwc.dtype != torch.bfloat16:
            nhwc = nhwc.float()
        return self.levels(self.launch(nhwc.reshape(V * H * W, C), V, H, W), V, H, W)
