"""Drop-in for ``ldm.models.autoencoder.VQModelInterface`` (reference autoencoder.py:14-110, 264-282), decode side.

``decode(h, force_not_quantize=False)`` = VQ nearest-code lookup (taming VectorQuantizer2, restated as the
stedm_vq_nearest kernel) -> post_quant_conv -> Decoder, all on the native engine.  Parameter names follow the
reference (``quantize.embedding.weight``, ``post_quant_conv``, ``quant_conv``, ``decoder.*``) so vq-f4.ckpt loads.
The encoder (used by the reference's get_input only to produce a tensor whose length is read) is
outside the sampling path: encoder.* keys in a checkpoint are accepted and ignored.
"""
import torch
import torch.nn as nn

from .. import util as ldm_util  # noqa: F401
from ..modules.diffusionmodules.model import Decoder


class VectorQuantizer(nn.Module):
    """Holds the codebook under the reference's name; lookup runs in ops.vq_nearest."""

    def __init__(self, n_e, e_dim, beta=0.25):
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    @torch.no_grad()
    def forward(self, z):
        from ... import ops
        zq, idx = ops.vq_nearest(z.float().contiguous(), self.embedding.weight.detach().float().contiguous(), True)
        return zq, None, (None, None, idx.long())

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q


class VQModelInterface(nn.Module):
    def __init__(self, embed_dim, ddconfig=None, lossconfig=None, n_embed=8192, ckpt_path=None, ignore_keys=(),
                 image_key="image", colorize_nlabels=None, monitor=None, batch_resize_range=None,
                 scheduler_config=None, lr_g_factor=1.0, remap=None, sane_index_shape=False, use_ema=False,
                 precision="bf16"):
        super().__init__()
        assert remap is None and not sane_index_shape and not use_ema
        dd = dict(ddconfig)
        self.embed_dim, self.n_embed, self.image_key = embed_dim, n_embed, image_key
        self.decoder = Decoder(**dd)
        self.quantize = VectorQuantizer(n_embed, embed_dim, beta=0.25)
        self.quant_conv = nn.Conv2d(dd["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, dd["z_channels"], 1)
        self.precision = precision
        self._runner = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=list(ignore_keys))

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        missing, unexpected = self.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    # packed-weight lifecycle: same contract as UNetModel (see its comment)
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

    def _weights_signature(self):
        return tuple((p._version, p.data_ptr()) for p in self.parameters())

    def runner(self):
        if self._runner is not None and self._runner_sig != self._weights_signature():
            self._runner = None
        if self._runner is None:
            from ...engine import DecoderRunner
            if not self.post_quant_conv.weight.is_cuda:
                raise RuntimeError("VQModelInterface.decode runs only on a CUDA (sm_100a) device")
            self._runner = DecoderRunner(self, self.precision)
            self._runner_sig = self._weights_signature()
        return self._runner

    def encode(self, x):
        raise NotImplementedError("the first-stage encoder is outside the sampling path (SURVEY.md §8f rank 1)")

    @torch.no_grad()
    def decode(self, h, force_not_quantize=False):
        return self.runner()(h, force_not_quantize=force_not_quantize)
