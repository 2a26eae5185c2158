"""Drop-in for ``ldm.modules.encoders.modules.SpatialRescaler`` (reference encoders/modules.py:104-133): the layout
conditioner — n_stages bilinear x0.5 resizes followed by a bias-free 1x1 conv — as one fused kernel
(stedm_spatial_rescale)."""
import torch
import torch.nn as nn


class SpatialRescaler(nn.Module):
    def __init__(self, n_stages=1, method="bilinear", multiplier=0.5, in_channels=3, out_channels=None, bias=False):
        super().__init__()
        assert n_stages >= 0
        if method != "bilinear" or multiplier != 0.5 or bias or out_channels is None:
            raise NotImplementedError("SpatialRescaler options outside STEDM's cond_stage_config (spatial.yaml)")
        self.n_stages, self.multiplier = n_stages, multiplier
        self.remap_output = True
        self.channel_mapper = nn.Conv2d(in_channels, out_channels, 1, bias=False)

    @torch.no_grad()
    def forward(self, x):
        from .... import ops
        w = self.channel_mapper.weight.detach().float().reshape(self.channel_mapper.out_channels, -1).contiguous()
        return ops.spatial_rescale(x.float().contiguous(), w, self.n_stages)

    def encode(self, x):
        return self(x)
