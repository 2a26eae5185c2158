"""Style aggregation blocks, mirroring networks/agg_blocks.py of the reference (Agg_Linear :6-35, Agg_Max :38-54,
Agg_Mean :57-75, Agg_None :78-85) with the same registered names (``_embedder`` and ``embedder``).

The embedder module (torchvision ``swin_v2_t`` with a 768->512 head, s_zss_dm.py:19-20) only OWNS the parameters, so
reference checkpoints load by name; its forward is executed natively by stedm_b200.style_engine.StyleEncoderRunner
(patch embedding, window attention, LayerNorm kernels + tcgen05 GEMMs) straight from the ``b n h w c`` style tensor —
the reference's ``(b n) c h w`` permute is folded into the patch-embedding kernel — and the reduction over the n
style images of a sample is stedm_set_reduce.  No CPU or torchvision fallback: CUDA tensors only.
"""
import torch

from .. import ops


class _Agg(torch.nn.Module):
    def __init__(self, sampling_cfg, embedder):
        super().__init__()
        self._sampling_cfg = sampling_cfg
        self._embedder = embedder
        self.register_module("embedder", self._embedder)
        self.precision = "bf16"
        self._runner = None

    # ---- packed-weight lifecycle (same contract as UNetModel / VQModelInterface) -------------------------
    def invalidate_packed(self):
        self._runner = None

    def _apply(self, fn, *a, **k):
        self._runner = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._runner = None
        return super().load_state_dict(*a, **k)

    def set_precision(self, precision):
        if precision != self.precision:
            self.precision, self._runner = precision, None

    def _versions(self):
        return tuple(p._version for p in self._embedder.parameters())

    def runner(self):
        ver = self._versions()
        if self._runner is None or self._runner_versions != ver:
            from ..style_engine import StyleEncoderRunner
            if not next(self._embedder.parameters()).is_cuda:
                raise RuntimeError("the style encoder runs only on a CUDA (sm_100a) device: move the model with "
                                   ".cuda() first — there is no CPU path")
            self._runner = StyleEncoderRunner(self._embedder, self.precision)
            self._runner_versions = ver
        return self._runner

    def embed(self, style_imgs):
        """'b n h w c' style images -> per-image features [b, n, 512] (agg_blocks.py:26-30, 49-52, 68-72)."""
        b, n, h, w, c = style_imgs.shape
        if not style_imgs.is_cuda:
            raise RuntimeError("stedm_b200 style encoder takes CUDA tensors only (no CPU fallback)")
        feats = self.runner()(style_imgs.reshape(b * n, h, w, c))
        return feats.view(b, n, -1)


class Agg_Mean(_Agg):
    def forward(self, style_imgs):
        return ops.set_reduce(self.embed(style_imgs), "mean")


class Agg_Max(_Agg):
    def forward(self, style_imgs):
        return ops.set_reduce(self.embed(style_imgs), "max")


class Agg_Linear(_Agg):
    def __init__(self, sampling_cfg, embedder):
        super().__init__(sampling_cfg, embedder)
        num = sampling_cfg.num_patches if sampling_cfg.name == "mp" else 1
        self._linear_block = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Linear(512 * num, 512), torch.nn.ReLU(),
                                                 torch.nn.Linear(512, 512), torch.nn.ReLU())
        self.register_module("linear_block", self._linear_block)

    def forward(self, style_imgs):
        f = self.embed(style_imgs)
        f = f.reshape(f.shape[0], -1)
        l1, l3 = self._linear_block[1], self._linear_block[3]
        # ReLU -> Linear -> ReLU -> Linear -> ReLU (agg_blocks.py:15-19) on the K8 linear kernel
        h = ops.linear(f, l1.weight.detach().float().contiguous(), l1.bias.detach().float().contiguous(), act_in="relu")
        return ops.linear(h, l3.weight.detach().float().contiguous(), l3.bias.detach().float().contiguous(),
                          act_in="relu", relu_out=True)


class Agg_None(torch.nn.Module):
    def __init__(self, sampling_cfg, embedder):
        super().__init__()
        self._sampling_cfg = sampling_cfg
        self._embedder = embedder

    def forward(self, style_imgs):
        return torch.zeros((style_imgs.shape[0], 512), dtype=style_imgs.dtype, device=style_imgs.device)
