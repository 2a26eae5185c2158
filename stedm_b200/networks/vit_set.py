"""Drop-in for ``networks.vit_set.sViT`` (reference vit_set.py:109-208), the ``style_agg=svit`` aggregator built by
networks/s_zss_dm.py:31-38: a ViT over the SET of style images of a sample (images stacked along channels, SPT patch
tokens, cls + time tokens, LSA attention with a learned temperature and a masked diagonal, mean/cls pooling, MLP head).

The modules below only OWN parameters under the reference's names (``to_patch_embedding.to_patch_tokens.{1,2}``,
``pos_embedding``, ``cls_token``, ``transformer.layers.N.0.{norm,fn.temperature,fn.to_qkv,fn.to_out.0}``,
``transformer.layers.N.1.{norm,fn.net.0,fn.net.3}``, ``mlp_head.{0,1}``, ``to_time_embedding``), so reference
checkpoints load unchanged; the forward pass is executed by stedm_b200.style_engine.SetViTRunner on the C-ABI kernels.
Inference only (dropout layers are identities); CUDA tensors only.
"""
import torch
from torch import nn


def _pair(t):
    return t if isinstance(t, tuple) else (t, t)


class _PreNorm(nn.Module):          # vit_set.py:15-21
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _FeedForward(nn.Module):      # vit_set.py:23-34: Linear, GELU, Dropout, Linear, Dropout
    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Identity(), nn.Linear(hidden_dim, dim),
                                 nn.Identity())


class _LSA(nn.Module):              # vit_set.py:36-65
    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head = heads, dim_head
        self.temperature = nn.Parameter(torch.log(torch.tensor(dim_head ** -0.5)))
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Identity())


class _Transformer(nn.Module):      # vit_set.py:67-80
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([_PreNorm(dim, _LSA(dim, heads, dim_head)),
                                                    _PreNorm(dim, _FeedForward(dim, mlp_dim))]) for _ in range(depth)])


class _SPT(nn.Module):              # vit_set.py:82-107
    def __init__(self, dim, patch_size, channels, sample_size):
        super().__init__()
        patch_dim = patch_size * patch_size * sample_size * channels
        self.to_patch_tokens = nn.Sequential(nn.Identity(), nn.LayerNorm(patch_dim), nn.Linear(patch_dim, dim))


class sViT(nn.Module):
    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool="cls", channels=3,
                 dim_head=64, dropout=0., emb_dropout=0., ns=5, t_dim=256):
        super().__init__()
        ih, iw = _pair(image_size)
        ph, pw = _pair(patch_size)
        assert ih % ph == 0 and iw % pw == 0, "Image dimensions must be divisible by the patch size."
        assert pool in {"cls", "mean", "none"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        if ih != iw or ph != pw or channels != 3:
            raise NotImplementedError("native sViT: square images / patches with 3 channels")
        self.ns, self.np = ns, (ih // ph) * (iw // pw)
        self.image_size, self.patch_size, self.pool = ih, ph, pool
        self.to_patch_embedding = _SPT(dim, ph, channels, ns)
        self.pos_embedding = nn.Parameter(torch.randn(1, self.np + 2, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.transformer = _Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self.to_time_embedding = nn.Linear(t_dim, dim)
        self.precision = "bf16"
        self._runner = None

    # ---- packed-weight lifecycle --------------------------------------------------------------------------
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

    def runner(self):
        ver = tuple(p._version for p in self.parameters())
        if self._runner is None or self._runner_versions != ver:
            from ..style_engine import SetViTRunner
            if not self.cls_token.is_cuda:
                raise RuntimeError("sViT runs only on a CUDA (sm_100a) device: move the model with .cuda() first — "
                                   "there is no CPU path")
            self._runner, self._runner_versions = SetViTRunner(self, self.precision), ver
        return self._runner

    @torch.no_grad()
    def forward(self, img, t_emb=None, c_old=None):
        """img: 'b n h w c' style images (vit_set.py:163-208).  The sampling path calls it with the images only
        (s_zss_dm.py:55); the time-token / iterative-conditioning inputs are training-time options."""
        if t_emb is not None or c_old is not None:
            raise NotImplementedError("native sViT: t_emb / c_old are not used on the sampling path")
        if not img.is_cuda:
            raise RuntimeError("stedm_b200 sViT takes CUDA tensors only (no CPU fallback)")
        return self.runner()(img)
