"""Drop-in for ``ldm.modules.diffusionmodules.openaimodel.UNetModel`` (reference openaimodel.py:435-806).

Same constructor parameters (conf/diffusion/unet_config/landscape.yaml), same parameter names and shapes
 so reference checkpoints load with ``load_state_dict``; ``forward(x, timesteps, context)`` has
the reference's signature and NCHW fp32 tensors in/out — but it executes on the native sm_100a engine
(stedm_b200.engine.UNetRunner).  The nn modules below only OWN parameters; their own ``forward`` is never used.

Structure executed by the reference with the shipped config: a stem conv, per
level ``num_res_blocks`` ResBlocks and a strided-conv Downsample between levels; middle = ResBlock,
ResBlockStyle (style vector in place of the timestep embedding), AttentionBlock (legacy head-major qkv),
ResBlock; decoder mirrors it with skip concatenation and nearest-x2 + conv Upsamples; GN-SiLU-conv head.
No attention block exists outside the middle block: the reference's ``ds in attention_resolutions`` branches
are dead (``layers.append()`` with its argument commented out raises TypeError, openaimodel.py:580-590), so a
config that would reach them is rejected here too.
"""
import torch
import torch.nn as nn


def _norm(ch):
    return nn.GroupNorm(32, ch)  # GroupNorm32: eps 1e-5, fp32 statistics (util.py:199-216)


class ResBlock(nn.Module):
    """Parameter container for openaimodel.py:176-288 (use_scale_shift_norm=False, no up/down)."""

    def __init__(self, channels, emb_channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(_norm(channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(_norm(self.out_channels), nn.SiLU(), nn.Dropout(p=0.0),
                                        nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)
        # reference: zero_module on the second conv (openaimodel.py:242-244)
        nn.init.zeros_(self.out_layers[3].weight)
        nn.init.zeros_(self.out_layers[3].bias)


class ResBlockStyle(nn.Module):
    """openaimodel.py:291-297: a ResBlock whose embedding input is the style vector (context)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.block = ResBlock(*args, **kwargs)


class AttentionBlock(nn.Module):
    """openaimodel.py:300-346; qkv/proj are Conv1d k=1; QKVAttentionLegacy ordering."""

    def __init__(self, channels, num_heads):
        super().__init__()
        self.channels, self.num_heads = channels, num_heads
        self.norm = _norm(channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = nn.Conv1d(channels, channels, 1)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)


class Downsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.op = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=1)  # openaimodel.py:164-166


class Upsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)          # openaimodel.py:120


class _Seq(nn.Sequential):
    """Stand-in for TimestepEmbedSequential (index-compatible parameter names); ``kind`` tags the role."""
    kind = "res"


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 style_imgs=1, precision="bf16"):
        super().__init__()
        unsupported = dict(dropout=dropout != 0, dims=dims != 2, num_classes=num_classes is not None,
                           use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                           use_new_attention_order=use_new_attention_order, conv_resample=not conv_resample,
                           n_embed=n_embed is not None, num_head_channels=num_head_channels != -1)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"UNetModel options outside STEDM's sampling path: {bad}")
        if use_spatial_transformer:                                         # openaimodel.py:494-501
            assert context_dim is not None, "Fool!! You forgot to include the dimension of your cross-attention conditioning..."
        if context_dim is not None:
            assert use_spatial_transformer, "Fool!! You forgot to use the spatial transformer for your cross-attention conditioning..."
            if not isinstance(context_dim, int):
                context_dim = list(context_dim)
        if num_heads == -1:
            raise AssertionError("Either num_heads or num_head_channels has to be set")
        channel_mult = tuple(channel_mult)
        ds = 1
        for level in range(len(channel_mult)):
            if ds in list(attention_resolutions):
                # the reference constructor raises TypeError here (list.append() without an argument)
                raise TypeError(f"attention_resolutions contains downsample rate {ds}: the reference cannot build "
                                f"this configuration (openaimodel.py:580-590)")
            if level != len(channel_mult) - 1:
                ds *= 2
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks, self.channel_mult = out_channels, num_res_blocks, channel_mult
        self.attention_resolutions = list(attention_resolutions)
        self.num_heads = num_heads
        self.dtype = torch.float32
        self.precision = precision
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))

        stem = _Seq(nn.Conv2d(in_channels, model_channels, 3, padding=1))
        stem.kind = "stem"
        self.input_blocks = nn.ModuleList([stem])
        skip_chans = [model_channels]
        ch = model_channels
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                self.input_blocks.append(_Seq(ResBlock(ch, ted, mult * model_channels)))
                ch = mult * model_channels
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                down = _Seq(Downsample(ch))
                down.kind = "down"
                self.input_blocks.append(down)
                skip_chans.append(ch)
        if use_spatial_transformer:
            # middle_block[2] = SpatialTransformer(ch, num_heads, ch // num_heads, depth, context_dim)
            # (openaimodel.py:625-652, legacy=True); it is called without a context (see engine.UNetRunner)
            from ..attention import SpatialTransformer
            mid_attn = SpatialTransformer(ch, num_heads, ch // num_heads, depth=transformer_depth,
                                          context_dim=context_dim, precision=precision)
        else:
            mid_attn = AttentionBlock(ch, num_heads)
        self.middle_block = _Seq(ResBlock(ch, ted), ResBlockStyle(ch, ted), mid_attn, ResBlock(ch, ted))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + skip_chans.pop(), ted, model_channels * mult)]
                ch = model_channels * mult
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                self.output_blocks.append(_Seq(*layers))
        self.out = nn.Sequential(_norm(ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))
        nn.init.zeros_(self.out[2].weight)
        nn.init.zeros_(self.out[2].bias)
        self._runner = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # ---- packed-weight lifecycle ------------------------------------------------------------------------
    # The runner holds derived copies of the weights (bf16 repacks, stacked embedding tables, the per-timestep embedding
    # cache) and captured CUDA graphs.  They are dropped whenever the nn.Parameters can have changed: .to()/.cuda()
    # (_apply), load_state_dict on THIS module or on any parent (the post hook fires during a parent's recursive load,
    # which never calls a child's load_state_dict override), and in-place updates that bump a parameter's version
    # counter or move its storage (checked by runner()).  Writes through ``p.data`` bypass the version counter: call
    # invalidate_packed() after those.
    def invalidate_packed(self):
        self._runner = None
        self.__dict__.pop("_graph_cache", None)  # captured graphs hold the old packed weights

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    def set_precision(self, precision):
        if precision != self.precision:
            self.precision = precision
            self.invalidate_packed()

    def _weights_signature(self):
        return tuple((p._version, p.data_ptr()) for p in self.parameters())

    def runner(self):
        if self._runner is not None and self._runner_sig != self._weights_signature():
            self.invalidate_packed()
        if self._runner is None:
            from ....engine import UNetRunner
            if not next(self.parameters()).is_cuda:
                raise RuntimeError("UNetModel runs only on a CUDA (sm_100a) device: move the model with .cuda() "
                                   "first — there is no CPU path")
            self._runner = UNetRunner(self, self.precision)
            self._runner_sig = self._weights_signature()
        return self._runner

    # ---- reference API ----------------------------------------------------------------------------------
    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """x: (B, in_channels, H, W) NCHW fp32 (already concatenated with c_concat, ddpm.py:1415),
        timesteps (B,), context (B, 4*model_channels) -> eps (B, out_channels, H, W) fp32."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise RuntimeError("the native U-Net is inference-only (no autograd)")
        return self.runner()(x.float(), None, timesteps.to(torch.int64), context)

    def forward_split(self, x, c_concat, timesteps, context, uniform_t=False, emb=None):
        """Same as forward(cat([x, c_concat], 1), ...) with the concat fused into the input packing kernel.
        ``context`` may hold G*B rows for B inputs (guided sampling with a shared encoder trunk, see
        engine.UNetRunner.__call__); eps then has G*B rows."""
        return self.runner()(x.float(), c_concat.float().contiguous(), timesteps.to(torch.int64), context,
                             uniform_t=uniform_t, emb=emb)

    def shared_trunk_ok(self, batch, latent_hw):
        return self.runner().shared_trunk_ok(batch, latent_hw)
