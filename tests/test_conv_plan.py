"""Host-side launch planning of the tensor-core convolution (stedm_conv_tc_plan): tile geometry checks, channel tile /
CTA-pair choice, halo-mode selection and ring sizing.  Pure host code — runs without a GPU; the kernels that execute
these plans are checked by tests/test_gpu_tc.py."""
import ctypes as C

import pytest

from stedm_b200 import _lib

BF16 = _lib.BF16
DUMMY = 4096  # a non-NULL "pointer": the plan never dereferences it


def plan(B, H, W, c0, cout, k, *, c1=0, skip=0, tap_mode=0, stats=False, workspace=0, x1_batch=0):
    d = _lib.ConvDesc()
    d.x0, d.weight, d.out = DUMMY, DUMMY, DUMMY
    d.x1 = DUMMY if c1 else None
    d.batch, d.in_h, d.in_w, d.c0, d.c1, d.cout, d.ksize = B, H, W, c0, c1, cout, k
    d.in_dtype, d.out_dtype, d.stride, d.upsample = BF16, BF16, 1, 0
    d.tap_mode, d.phase, d.x1_batch = tap_mode, 0, x1_batch
    if skip:
        d.skip_x0, d.skip_c0 = DUMMY, skip
    if stats:
        d.stats_out = DUMMY
    if workspace:
        d.workspace, d.workspace_bytes = DUMMY, workspace
    out = (C.c_int32 * 8)()
    rc = _lib.load().stedm_conv_tc_plan(C.byref(d), out)
    if rc != 0:
        raise RuntimeError(_lib.last_error())
    return dict(zip(("bn", "cl", "pair", "halo", "slots", "slot_bytes", "ksplit", "k_slabs"), out))


def test_bench_shapes_pick_cta_pairs_and_halo_boxes():
    # 1024 -> 1024 3x3 at 16x16 on 128 samples: BN 256 CTA pairs; box = (8 + 2) rows x 16 px x 128 B, 4 of them in 96 KB
    p = plan(128, 16, 16, 1024, 1024, 3, stats=True)
    assert p == dict(bn=256, cl=2, pair=1, halo=1, slots=4, slot_bytes=10 * 16 * 128, ksplit=1, k_slabs=144)
    # 32x32: (4 + 2) rows x 32 px = 24 KB boxes
    p = plan(128, 32, 32, 512, 512, 3, stats=True)
    assert (p["halo"], p["slots"], p["slot_bytes"]) == (1, 4, 6 * 32 * 128)
    # 64x64, 128 output channels: BN 128 pairs with a 64 KB pool -> two 32 KB boxes
    p = plan(64, 64, 64, 128, 128, 3, stats=True)
    assert (p["bn"], p["pair"], p["halo"], p["slots"], p["slot_bytes"]) == (128, 1, 1, 2, 4 * 64 * 128)
    # the 16-channel eps head
    p = plan(128, 64, 64, 128, 16, 3)
    assert (p["bn"], p["cl"], p["halo"], p["slots"]) == (16, 1, 1, 2)
    # sub-pixel phase of the folded upsample: 2x2 taps, one halo row
    p = plan(128, 16, 16, 1024, 1024, 3, tap_mode=1, stats=True)
    assert (p["halo"], p["slot_bytes"], p["k_slabs"]) == (1, 9 * 16 * 128, 64)


def test_one_slab_per_tap_cases():
    assert plan(128, 16, 16, 1024, 3072, 1)["halo"] == 0                 # 1x1: nothing to share between taps
    assert plan(64, 128, 128, 128, 128, 3, stats=True)["halo"] == 0      # W = 128: tiles are single image rows
    assert plan(3, 8, 8, 64, 128, 3)["halo"] == 0                        # two samples per tile
    assert plan(1, 8, 16, 64, 64, 3)["halo"] == 0                        # H < tile rows + halo
    # fused-skip launches: halo only while the skip slabs stay below a fifth of the tap slabs
    assert plan(128, 16, 16, 1024, 1024, 3, skip=1536, stats=True)["halo"] == 1      # 24 * 5 <= 144
    assert plan(128, 32, 32, 512, 512, 3, skip=1536, stats=True)["halo"] == 0        # 24 * 5 > 72
    assert plan(128, 64, 64, 128, 128, 3, skip=640, stats=True)["halo"] == 0
    p = plan(128, 16, 16, 1024, 3072, 1)
    assert (p["slots"], p["slot_bytes"]) == (6, 128 * 128)               # lock-step rings: 6 x 16 KB


def test_split_k_needs_a_workspace_and_disables_halo():
    small = dict(B=1, H=16, W=16, c0=1024, cout=1024, k=3)
    assert plan(**small)["ksplit"] == 1 and plan(**small)["halo"] == 1   # no workspace: single pass
    p = plan(**small, workspace=1 << 30)
    assert p["ksplit"] > 1 and p["halo"] == 0
    # fused GroupNorm statistics no longer force the single pass: the split-K finish pass publishes them (round 2)
    assert plan(**small, workspace=1 << 30, stats=True)["ksplit"] > 1
    # a launch whose tiles already fill the grid is never split
    assert plan(B=64, H=16, W=16, c0=1024, cout=1024, k=3, workspace=1 << 30)["ksplit"] == 1


@pytest.mark.parametrize("kw,msg", [
    (dict(B=1, H=16, W=96, c0=64, cout=64, k=3), "width"),               # 96 neither divides nor is a multiple of 128
    (dict(B=1, H=6, W=32, c0=64, cout=64, k=3), "height"),               # 4-row tiles do not tile 6 rows
    (dict(B=1, H=16, W=16, c0=72, cout=64, k=3), "channel counts"),      # 3x3 needs whole 64-channel slabs
    (dict(B=1, H=16, W=16, c0=64, cout=24, k=3), "cout"),
])
def test_rejected_geometry_reports_why(kw, msg):
    with pytest.raises(RuntimeError, match=msg):
        plan(**kw)


def test_plain_gemm_relaxations():
    # token-major linears of the style encoder: K = 96 (partial slab, zero-filled by TMA), N = 288 (partial last tile)
    # (the channel tile with the least padding wins: 288 -> 5 x 64, 96 -> 1 x 128)
    p = plan(1, 1, 4096 * 64, 96, 288, 1)
    assert (p["bn"], p["k_slabs"], p["halo"]) == (64, 2, 0)
    assert plan(1, 1, 4096 * 64, 96, 96, 1)["bn"] == 128
