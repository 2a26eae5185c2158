"""Deterministic random-init weights ("fixture weights") keyed by parameter NAME.

The reference's constructors zero-initialise every ResBlock's second conv, the attention ``proj_out`` and
the final ``out`` conv (openaimodel.py:242-244, 334, 732) so a freshly built U-Net predicts eps == 0 and the
guidance rescale divides 0/0; the VQ codebook is U(-1/8192, 1/8192) so every latent quantises to ~0.  For
parity and for benchmarking on "random-init weights of that architecture" every floating point ``.weight`` /
``.bias`` is therefore re-drawn from a generator seeded by ``crc32(canonical name) ^ seed``.  Because the
draw depends only on the name and shape, the reference model (``oracle/make_golden.py``), the CPU oracle and
the CUDA engine all receive bit-identical tensors without shipping a 1 GB checkpoint.
"""
import zlib

import torch

CODEBOOK_STD = 64.0  # matched to the std of final latents under fixture weights (see make_golden log)


def canonical_name(key: str) -> str:
    """Collapse the reference's duplicated registrations (``_agg_block.`` / ``agg_block.``, ``._embedder.`` /
    ``.embedder.``, and LDM_Diffusion's ``_model.`` / ``model.`` prefixes) onto one spelling."""
    k = key
    if k.startswith("_model."):
        k = "model." + k[len("_model."):]
    k = k.replace("_agg_block.", "agg_block.")
    k = k.replace("._embedder.", ".embedder.")
    return k


def fixture_tensor(key: str, shape, seed: int = 0, codebook_std: float = CODEBOOK_STD) -> torch.Tensor:
    name = canonical_name(key)
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    t = torch.randn(shape, generator=g, dtype=torch.float32)
    if name.endswith("quantize.embedding.weight"):
        return t * codebook_std
    if name.endswith(".bias"):
        return t * 0.05
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return t * (1.0 / fan_in ** 0.5)
    return 1.0 + 0.1 * t  # 1-D ``.weight`` = normalisation gain


def is_fixture_key(key: str, tensor: torch.Tensor) -> bool:
    if not tensor.is_floating_point():
        return False
    if key.startswith("model_ema.") or ".model_ema." in key:
        return False
    # sViT's free-standing parameters (vit_set.py:132-133) are torch.randn in the constructor: name-keyed too
    return key.endswith((".weight", ".bias", "pos_embedding", "cls_token"))


@torch.no_grad()
def apply_fixture_weights(module: torch.nn.Module, seed: int = 0, codebook_std: float = CODEBOOK_STD):
    """Overwrite every floating point ``.weight``/``.bias`` of ``module`` in place (CPU draw, then copy)."""
    sd = module.state_dict()
    done = {}
    for key, tensor in sd.items():
        if not is_fixture_key(key, tensor):
            continue
        ptr = tensor.data_ptr()
        if ptr in done:
            continue
        done[ptr] = key
        tensor.copy_(fixture_tensor(key, tensor.shape, seed, codebook_std).to(tensor.dtype))
    return module
