"""Style aggregation blocks, mirroring networks/agg_blocks.py of the reference (Agg_Linear :6-35, Agg_Max :38-54,
Agg_Mean :57-75, Agg_None :78-85) with the same registered names (``_embedder`` and ``embedder``).

The style encoder itself is torchvision's ``swin_v2_t`` (a library call in the reference too, s_zss_dm.py:19-20);
it runs once per conditioning set, outside the DDIM loop (~0.1 % of the path's FLOPs) and is a
"next" row of the scope table.  The (b n) h w c -> c h w shuffles and the reductions over n are plain torch views.
"""
import torch


def _embed(embedder, style_imgs):
    b, n, h, w, c = style_imgs.shape
    imgs = style_imgs.permute(0, 1, 4, 2, 3).reshape(b * n, c, h, w)
    return embedder(imgs).reshape(b, n, -1)


class _Agg(torch.nn.Module):
    def __init__(self, sampling_cfg, embedder):
        super().__init__()
        self._sampling_cfg = sampling_cfg
        self._embedder = embedder
        self.register_module("embedder", self._embedder)


class Agg_Mean(_Agg):
    def forward(self, style_imgs):
        return torch.mean(_embed(self._embedder, style_imgs), dim=1)


class Agg_Max(_Agg):
    def forward(self, style_imgs):
        return torch.max(_embed(self._embedder, style_imgs), dim=1)[0]


class Agg_Linear(_Agg):
    def __init__(self, sampling_cfg, embedder):
        super().__init__(sampling_cfg, embedder)
        num = sampling_cfg.num_patches if sampling_cfg.name == "mp" else 1
        self._linear_block = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Linear(512 * num, 512), torch.nn.ReLU(),
                                                 torch.nn.Linear(512, 512), torch.nn.ReLU())
        self.register_module("linear_block", self._linear_block)

    def forward(self, style_imgs):
        f = _embed(self._embedder, style_imgs)
        return self._linear_block(f.reshape(f.shape[0], -1))


class Agg_None(torch.nn.Module):
    def __init__(self, sampling_cfg, embedder):
        super().__init__()
        self._sampling_cfg = sampling_cfg
        self._embedder = embedder

    def forward(self, style_imgs):
        return torch.zeros((style_imgs.shape[0], 512), dtype=style_imgs.dtype, device=style_imgs.device)
