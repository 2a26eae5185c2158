"""Host-side mirror of the reference interface (no GPU): checkpoint-name compatibility, the plugin seam, the
hydra-free config loader, fixture determinism, and the no-CPU-fallback rule."""
import json
import os

import pytest
import torch

from tests.util import build_config


@pytest.fixture(scope="module")
def model():
    from stedm_b200.modules.ldm_diffusion import LDM_Diffusion
    return LDM_Diffusion(build_config(32, n_style=1), load_first_stage_ckpt=False)


def test_state_dict_names_match_reference(model, golden_dir):
    """Every tensor of the reference's S_ZSS_DM state dict that is on the sampling path exists here with the same
    name and shape (key list dumped from the reference by oracle/make_golden tooling)."""
    ref = json.load(open(os.path.join(golden_dir, "reference_state_dict_keys.json")))
    mine = {k: list(v.shape) for k, v in model._model.state_dict().items()}
    skipped = ("model_ema.", "first_stage_model.encoder.")      # EMA shadow + VAE encoder: outside the path
    for k, shape in ref.items():
        if k.startswith(skipped):
            continue
        assert k in mine, k
        assert mine[k] == shape, (k, mine[k], shape)
    assert not [k for k in mine if k not in ref]


def test_reference_checkpoint_loads_non_strict(model, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "reference_state_dict_keys.json")))
    sd = {k: torch.zeros(s) for k, s in ref.items() if not k.endswith("relative_position_index")}
    missing, unexpected = model._model.load_state_dict(sd, strict=False)
    assert all(k.startswith(("model_ema.", "first_stage_model.encoder.")) for k in unexpected)
    assert all(k.endswith("relative_position_index") for k in missing)


def test_plugin_seam_redirects_reference_targets():
    from stedm_b200.ldm.util import get_obj_from_str, instantiate_from_config
    from stedm_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    from stedm_b200.ldm.modules.encoders.modules import SpatialRescaler
    assert get_obj_from_str("ldm.modules.diffusionmodules.openaimodel.UNetModel") is UNetModel
    assert get_obj_from_str("torch.nn.Identity") is torch.nn.Identity
    m = instantiate_from_config({"target": "ldm.modules.encoders.modules.SpatialRescaler",
                                 "params": {"n_stages": 2, "in_channels": 2, "out_channels": 3}})
    assert isinstance(m, SpatialRescaler) and tuple(m.channel_mapper.weight.shape) == (3, 2, 1, 1)
    with pytest.raises(KeyError):
        instantiate_from_config({"params": {}})


def test_unet_rejects_configs_the_reference_cannot_build():
    from stedm_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with pytest.raises(TypeError):      # ds=1 in attention_resolutions -> reference: list.append() TypeError
        UNetModel(64, 6, 64, 3, 1, [1], channel_mult=(1, 2), num_heads=4)
    with pytest.raises(AssertionError):  # openaimodel.py:494-498: the two spatial-transformer options come together
        UNetModel(64, 6, 64, 3, 1, [32], channel_mult=(1, 2), num_heads=4, use_spatial_transformer=True)
    with pytest.raises(AssertionError):
        UNetModel(64, 6, 64, 3, 1, [32], channel_mult=(1, 2), num_heads=4, context_dim=512)
    # the configuration the reference CAN build: a SpatialTransformer in middle_block[2], reference parameter names
    m = UNetModel(64, 6, 64, 3, 1, [32], channel_mult=(1, 2), num_heads=4, use_spatial_transformer=True, context_dim=128)
    keys = set(m.state_dict())
    assert {"middle_block.2.proj_in.weight", "middle_block.2.transformer_blocks.0.attn2.to_k.weight",
            "middle_block.2.transformer_blocks.0.ff.net.0.proj.weight", "middle_block.2.proj_out.bias"} <= keys
    with pytest.raises(NotImplementedError):
        UNetModel(64, 6, 64, 3, 1, [32], channel_mult=(1, 2), num_heads=4, use_scale_shift_norm=True)


def test_config_loader_composes_defaults_and_overrides():
    from stedm_b200.config import load_config
    c = load_config([])
    assert c.ddim_steps == 128 and c.cfg_scale == 1.5 and c.style_agg.name == "linear" and c.diffusion.image_size == 128
    assert c.diffusion.unet_config.params.channel_mult == [1, 4, 8] and c.data.patch_size == 512
    c = load_config(["style_agg=mean", "style_sampling=mp", "location=cluster", "ddim_steps=50",
                     "diffusion.image_size=64", "+predict_dir=/tmp/x"])
    assert c.style_agg.name == "mean" and c.style_sampling.num_patches == 10 and c.location.n_gpus == 2
    assert c.ddim_steps == 50 and c.diffusion.image_size == 64 and c.predict_dir == "/tmp/x"


def test_fixture_weights_are_name_keyed_and_nonzero(model):
    from stedm_b200.utils.fixture import apply_fixture_weights, fixture_tensor
    apply_fixture_weights(model._model, seed=0)
    sd = model._model.state_dict()
    k = "model.diffusion_model.out.2.weight"                       # zero-initialised by the constructor
    assert float(sd[k].abs().max()) > 0
    assert torch.equal(sd[k], fixture_tensor(k, sd[k].shape, 0))
    a, b = sd["agg_block._embedder.head.weight"], sd["_agg_block._embedder.head.weight"]
    assert torch.equal(a, b)
    assert float(sd["first_stage_model.quantize.embedding.weight"].std()) > 30


def test_schedule_buffers_registered(model):
    m = model._model
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "posterior_variance",
                 "posterior_mean_coef1", "logvar"):
        assert getattr(m, name).shape == (1000,) and getattr(m, name).dtype == torch.float32
    assert m.num_timesteps == 1000 and m.parameterization == "eps"


def test_no_cpu_fallback(model):
    """The product path must fail loudly off-GPU: no PyTorch/CPU fallback for the kernels."""
    from stedm_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(1, 8), torch.zeros(4, 8), None)
    with pytest.raises(RuntimeError):
        model._model.model.diffusion_model(torch.zeros(1, 6, 32, 32), torch.zeros(1, dtype=torch.long),
                                           context=torch.zeros(1, 512))
    with pytest.raises(RuntimeError):
        model._model.first_stage_model.decode(torch.zeros(1, 3, 32, 32))


def test_prepare_batch_merges_multi_class_layouts(model):
    """CATCH / HER2 layouts have K > 2 classes (conf/data/catch.yaml: 6, her2.yaml: 8): prepare_batch sums classes
    1..K-1 into the foreground channel and keeps two channels, channels-last (modules/ldm_diffusion.py:51-60)."""
    g = torch.Generator().manual_seed(0)
    B, K, P, N = 2, 6, 16, 3
    lab = torch.randint(0, K, (B, P, P), generator=g)
    seg_oh = torch.nn.functional.one_hot(lab, K).permute(0, 3, 1, 2).float()
    img = torch.rand(B, 3, P, P, generator=g)
    style = torch.rand(B, N, 3, P, P, generator=g)
    out = model.prepare_batch((img, seg_oh.clone(), lab, style, torch.arange(B)))
    assert tuple(out["image"].shape) == (B, P, P, 3) and tuple(out["style_imgs"].shape) == (B, N, P, P, 3)
    assert tuple(out["segmentation"].shape) == (B, P, P, 2)
    assert torch.equal(out["segmentation"][..., 0], (lab == 0).float())
    assert torch.equal(out["segmentation"][..., 1], (lab != 0).float())
    assert torch.equal(out["style_imgs"][1, 2, :, :, 0], style[1, 2, 0])


def test_parent_level_load_and_in_place_updates_invalidate_packed_weights(model):
    """nn.Module.load_state_dict on a PARENT never calls a child's load_state_dict override, so the packed
    runners (bf16 repacks, embedding cache, CUDA graphs) are dropped by a load_state_dict post hook instead; in-place
    parameter updates are caught by the (version, data_ptr) signature runner() compares."""
    unet = model._model.model.diffusion_model
    vq = model._model.first_stage_model
    for mod in (unet, vq):
        mod._runner = object()                       # stands for a built runner (building one needs a GPU)
    unet.__dict__["_graph_cache"] = {"k": 1}
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(sd, strict=False)          # top-level load, as predict.run / init_from_ckpt on the parent do
    assert unet._runner is None and vq._runner is None and "_graph_cache" not in unet.__dict__
    sig = unet._weights_signature()
    with torch.no_grad():
        unet.time_embed[0].weight.mul_(1.0)          # EMA-style in-place update: bumps the version counter
    assert unet._weights_signature() != sig
    sig = vq._weights_signature()
    with torch.no_grad():
        vq.quantize.embedding.weight.add_(0.0)
    assert vq._weights_signature() != sig


def test_predict_checkpoint_resolution(tmp_path):
    """predict.resolve_checkpoint: +ckpt_path (documented key), the reference's ckpt_name rule (predict_diff.py:39-44),
    hard errors for named-but-missing files, None when nothing was named."""
    from stedm_b200.config import load_config
    from stedm_b200.predict import resolve_checkpoint
    f = tmp_path / "x.ckpt"
    f.write_bytes(b"0")
    assert resolve_checkpoint(load_config([f"+ckpt_path={f}"])) == str(f)
    with pytest.raises(FileNotFoundError):
        resolve_checkpoint(load_config([f"+ckpt_path={tmp_path}/nope.ckpt"]))
    (tmp_path / "checkpoints").mkdir()
    (tmp_path / "checkpoints" / "mine.ckpt").write_bytes(b"0")
    cfg = load_config([f"location.result_dir={tmp_path}", "+ckpt_name=mine.ckpt"])
    assert resolve_checkpoint(cfg) == str(tmp_path / "checkpoints" / "mine.ckpt")
    with pytest.raises(FileNotFoundError):
        resolve_checkpoint(load_config([f"location.result_dir={tmp_path}", "+ckpt_name=other.ckpt"]))
    assert resolve_checkpoint(load_config([f"location.result_dir={tmp_path}"])) is None


def test_bench_roofline_traffic_comes_from_the_committed_ncu_summary():
    """bench.py's roofline.traffic is parsed from profiles/rNN_ncu_full_conv_tc.csv (dram read + write of the dominant
    kernel's first captured launch), not a pasted literal."""
    import importlib.util
    import os
    from tests.util import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.ncu_traffic_bytes()
    assert t is not None and 1.0e8 < t < 2.5e8, t          # algorithmic bytes of that launch: 153 MB


def test_strong_scaling_split_keeps_per_sample_noise():
    """One global batch split over N ranks (bench.py's strong_scaling block): rank r's synthetic inputs are the global
    samples [r*B/N, (r+1)*B/N) — x_T is keyed by global sample index, so the union over ranks equals the 1-rank batch."""
    import importlib.util
    import os
    from tests.util import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    full = bench.synthetic_batch(8, 32, 1, 0)[3]
    parts = [bench.synthetic_batch(2, 32, 1, r * 2)[3] for r in range(4)]
    assert torch.equal(torch.cat(parts, 0), full)
