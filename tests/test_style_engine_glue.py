"""Host-side logic of the native style encoder on CPU: StyleEncoderRunner driven through the torch stand-in for the
kernel layer (tests/fake_ops.py) must reproduce torchvision's swin_v2_t — the library the reference calls
(networks/s_zss_dm.py:19-20).  Pins weight repacking (patch-embedding tap order, zeroed k bias, folded relative position
bias and logit scale), block order, res-post-norm wiring, patch-merging order and the pooled head.  The kernels
themselves are checked by tests/test_gpu_style.py (-m gpu)."""
import pytest
import torch
import torchvision

from tests import fake_ops
from tests.util import max_abs


def _swin(seed=0):
    torch.manual_seed(seed)
    m = torchvision.models.get_model("swin_v2_t")
    m.head = torch.nn.Linear(768, 512)
    with torch.no_grad():           # make every parameter matter: biases, norms and logit scales off their defaults
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif "norm" in n and n.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif n.endswith("logit_scale"):
                p.uniform_(1.0, 5.0)    # some above log(100) = 4.6: exercises the clamp
    return m.eval()


@pytest.fixture()
def patched(monkeypatch):
    from stedm_b200 import style_engine
    monkeypatch.setattr(style_engine, "ops", fake_ops)
    return style_engine


@pytest.mark.parametrize("p", [64, 128])
def test_style_encoder_runner_glue_fp32(patched, p):
    """p = 64 -> maps 16, 8, 4, 2 (padded windows in three stages); p = 128 -> 32, 16, 8, 4 (shifted + padded)."""
    m = _swin()
    torch.manual_seed(1)
    imgs = torch.rand(2, p, p, 3) * 2 - 1
    runner = patched.StyleEncoderRunner(m, "fp32")
    with torch.no_grad():
        want = m(imgs.permute(0, 3, 1, 2))
        got = runner(imgs)
    assert tuple(got.shape) == (2, 512)
    assert max_abs(got, want) < 2e-5 * max(1.0, float(want.abs().max())), max_abs(got, want)


def test_style_encoder_runner_glue_bf16(patched):
    m = _swin(2)
    torch.manual_seed(3)
    imgs = torch.rand(2, 128, 128, 3) * 2 - 1
    runner = patched.StyleEncoderRunner(m, "bf16")
    with torch.no_grad():
        want = m(imgs.permute(0, 3, 1, 2))
        got = runner(imgs)
    rel = max_abs(got, want) / float(want.abs().max())
    assert rel < 3e-2, rel


def test_relative_position_bias_matches_torchvision():
    from stedm_b200.style_engine import relative_position_bias
    m = _swin(4)
    at = m.features[3][1].attn
    with torch.no_grad():
        assert max_abs(relative_position_bias(at), at.get_relative_position_bias()[0]) < 1e-6


# ---------------------------------------------------------------------------------------------- style_agg=svit
SVIT_KW = dict(patch_size=8, num_classes=512, dim=256, depth=6, heads=12, mlp_dim=256, channels=3, dropout=0.1,
               emb_dropout=0.1, t_dim=256)
SVIT_CASES = {"svit_p64_n2_mean": (64, 2, "mean", 11), "svit_p128_n1_cls": (128, 1, "cls", 12)}   # as oracle/make_golden.py


def build_svit(P, ns, pool):
    """Product parameter container with name-keyed fixture weights (same tensors make_golden gave the reference)."""
    from stedm_b200.networks.vit_set import sViT
    from stedm_b200.utils.fixture import apply_fixture_weights
    holder = torch.nn.Module()
    holder.agg_block = sViT(image_size=P, ns=ns, pool=pool, **SVIT_KW).eval()
    apply_fixture_weights(holder, seed=0)
    return holder


@pytest.mark.parametrize("name", sorted(SVIT_CASES))
def test_svit_oracle_matches_reference_golden(name):
    """oracle.svit_aggregate on fixture weights == the reference's own sViT output (tests/golden/svit.npz)."""
    from oracle import stedm_oracle as O
    from tests.util import load_golden
    P, ns, pool, seed = SVIT_CASES[name]
    holder = build_svit(P, ns, pool)
    sd = {k: v.detach().float() for k, v in holder.state_dict().items()}
    _, style, _ = O.synthetic_batch(2, P, ns, seed)
    with torch.no_grad():
        got = O.svit_aggregate(sd, style, heads=12, patch=8, pool=pool)
    assert max_abs(got, load_golden("svit")[name]) < 1e-5


@pytest.mark.parametrize("name", sorted(SVIT_CASES))
@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_svit_runner_glue(patched, name, precision, bar):
    from oracle import stedm_oracle as O
    from tests.util import load_golden
    P, ns, pool, seed = SVIT_CASES[name]
    holder = build_svit(P, ns, pool)
    _, style, _ = O.synthetic_batch(2, P, ns, seed)
    with torch.no_grad():
        got = patched.SetViTRunner(holder.agg_block, precision)(style)
    want = load_golden("svit")[name]
    assert tuple(got.shape) == (2, 512)
    assert max_abs(got, want) < bar * (float(abs(want).max()) if precision == "bf16" else 1.0), max_abs(got, want)
