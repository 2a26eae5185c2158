"""Import shims that let the UNMODIFIED reference (/root/reference) run on a CPU-only box.

TEST INFRASTRUCTURE ONLY.  Nothing in ``stedm_b200/`` may import this module; it is used by
``oracle/make_golden.py`` (to generate ``tests/golden/*``) and by tests that are skipped when
``/root/reference`` is absent (it does not exist on the GPU box).

The reference depends on three packages that are not installed here and cannot be installed
(no network): ``pytorch_lightning``, ``taming`` and ``omegaconf``.  The shims below supply exactly
the names the sampling path touches:

* ``pytorch_lightning.LightningModule``  -> ``torch.nn.Module`` plus a ``device`` property and no-op
  ``log`` / ``save_hyperparameters`` (used by ldm/models/diffusion/ddpm.py:47, autoencoder.py:14).
* ``pytorch_lightning.utilities.rank_zero.rank_zero_only`` -> identity (ddpm.py:21).
* ``taming.modules.vqvae.quantize.VectorQuantizer2`` -> restatement of the published taming-transformers
  forward (CompVis/taming-transformers @master, taming/modules/vqvae/quantize.py, class VectorQuantizer2;
  the reference pins it only to ``@master`` in environment.yml).  Call sites: ldm/models/autoencoder.py:6,
  39-41, 277.  Used with beta=0.25, remap=None, sane_index_shape=False.
* ``omegaconf.listconfig.ListConfig`` -> empty class (openaimodel.py:503, only touched when context_dim set).

and one monkey patch: ``DDIMSampler.register_buffer`` hard-codes ``.to("cuda")`` (ldm/models/diffusion/
ddim.py:18-22), which fails without a GPU; it is replaced by a plain ``setattr``.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("STEDM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ldm"))


class _LightningModule(nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def save_hyperparameters(self, *a, **k):
        pass


class VectorQuantizer2(nn.Module):
    """Nearest-codebook-entry quantiser (restated from the published taming forward).

    z (b,c,h,w) -> (b,h,w,c) -> flat (N,c);  d = |z|^2 + |e|^2 - 2 z e^T;  idx = argmin_n d;
    z_q = E[idx]; straight-through z + (z_q - z).detach(); back to (b,c,h,w).
    """

    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random",
                 sane_index_shape=False, legacy=True):
        super().__init__()
        assert remap is None
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
        self.sane_index_shape = sane_index_shape

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        z = z.permute(0, 2, 3, 1).contiguous()
        zf = z.view(-1, self.e_dim)
        d = (torch.sum(zf ** 2, dim=1, keepdim=True)
             + torch.sum(self.embedding.weight ** 2, dim=1)
             - 2 * torch.einsum("bd,dn->bn", zf, self.embedding.weight.t()))
        idx = torch.argmin(d, dim=1)
        z_q = self.embedding(idx).view(z.shape)
        loss = None
        z_q = z + (z_q - z).detach()
        z_q = z_q.permute(0, 3, 1, 2).contiguous()
        return z_q, loss, (None, None, idx)

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q


def install():
    """Register the stub modules and put the reference on sys.path. Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        pl.LightningDataModule = object
        util = types.ModuleType("pytorch_lightning.utilities")
        rz = types.ModuleType("pytorch_lightning.utilities.rank_zero")
        rz.rank_zero_only = lambda f: f
        util.rank_zero = rz
        # older import path used by some files
        dist = types.ModuleType("pytorch_lightning.utilities.distributed")
        dist.rank_zero_only = lambda f: f
        util.distributed = dist
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
        sys.modules["pytorch_lightning.utilities.rank_zero"] = rz
        sys.modules["pytorch_lightning.utilities.distributed"] = dist
    if "taming" not in sys.modules:
        names = ["taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"]
        mods = [types.ModuleType(n) for n in names]
        for n, m in zip(names, mods):
            sys.modules[n] = m
        mods[0].modules = mods[1]
        mods[1].vqvae = mods[2]
        mods[2].quantize = mods[3]
        mods[3].VectorQuantizer2 = VectorQuantizer2
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        oc.ListConfig = ListConfig
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from ldm.models.diffusion.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)


class _NS(dict):
    """dict with attribute access — stands in for the hydra DictConfig nodes S_ZSS_DM reads."""
    __getattr__ = dict.__getitem__


def unet_config(image_size=128):
    return {"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel",
            "params": dict(image_size=image_size, in_channels=6, out_channels=3, model_channels=128,
                           attention_resolutions=[32, 16, 8], num_res_blocks=2, channel_mult=[1, 4, 8],
                           num_heads=8)}


def first_stage_config(resolution=512):
    return _NS({"target": "ldm.models.autoencoder.VQModelInterface",
                "params": _NS(embed_dim=3, n_embed=8192, monitor="Val Loss",
                              ddconfig=_NS(double_z=False, z_channels=3, resolution=resolution, in_channels=3,
                                           out_ch=3, ch=128, ch_mult=[1, 2, 4], num_res_blocks=2,
                                           attn_resolutions=[], dropout=0.0),
                              lossconfig={"target": "torch.nn.Identity"})})


def cond_stage_config():
    return {"target": "ldm.modules.encoders.modules.SpatialRescaler",
            "params": dict(n_stages=2, in_channels=2, out_channels=3)}


def build_reference_model(latent_size=64, style_sampling="augmented", style_agg="mean", num_patches=1):
    """Construct the reference's S_ZSS_DM exactly as LDM_Diffusion.__init__ does
    (modules/ldm_diffusion.py:27-38) minus the two checkpoint loads."""
    install()
    import contextlib
    import io
    from networks.s_zss_dm import S_ZSS_DM
    sampling = _NS(name=style_sampling, num_patches=num_patches)
    agg = _NS(name=style_agg)
    cfg = _NS(data=_NS(patch_size=latent_size * 4))
    with contextlib.redirect_stdout(io.StringIO()):
        model = S_ZSS_DM(
            encoder="swin_v2_t", sampling_cfg=sampling, agg_cfg=agg, cfg=cfg,
            first_stage_config=first_stage_config(latent_size * 4), cond_stage_config=cond_stage_config(),
            unet_config=unet_config(latent_size),
            beta_schedule="linear", linear_start=0.0015, linear_end=0.0205, num_timesteps_cond=1,
            log_every_t=100, timesteps=1000, loss_type="l1", first_stage_key="image",
            cond_stage_key="segmentation", image_size=latent_size, channels=3, conditioning_key="hybrid",
            cond_stage_trainable=True, monitor="Val Loss")
    model.eval()
    return model
