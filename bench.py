"""Benchmark of STEDM's synthetic-image sampling path on B200 (BASELINE.json: images/sec, DDIM-50, cfg 1.5, 256^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--latent L] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one generation batch: conditioning (layout rescaler + style encoder,
cond and uncond), the classifier-free-guided DDIM-50 loop over the eps U-Net (100 U-Net evaluations per image),
the first-stage VQ decode and the uint8 conversion — i.e. modules/ldm_diffusion.py:76-96 of the reference minus
PNG writing.  Workload at N=1 = BASELINE.json configs[1] (flowers, batch 64, 256^2, bf16); each extra rank gets its
own batch of 64 (weak scaling; per-sample noise keyed by global sample index; uint8 images all-gathered over NCCL).

Prints ONE JSON line (rank 0).  ``value`` = images/s with inputs resident in HBM; ``e2e`` = the same through
LDM_Diffusion.predict-style calls with pinned HOST buffers, H2D and D2H inside the timed region.
``--impl reference`` times the reference's algorithm on the host cores (oracle port — the reference itself is
Python and cannot be installed offline: pytorch_lightning / taming / omegaconf are absent) on a bounded sample.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

DDIM_STEPS, CFG_SCALE, ETA = 50, 1.5, 0.0
UNET_GFLOP_PER_SAMPLE_L64 = 217.29      # per sample per U-Net forward at latent 64
VAE_GFLOP_PER_SAMPLE_L64 = 670.6        # first-stage decode per image at 256^2
SWIN_GFLOP_PER_IMAGE_256 = 11.9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / power / throttle reasons sampled every 200 ms through NVML from a background thread while the
    timed region runs (NVML in-process: spawning nvidia-smi at 5 Hz was measured to slow the timed step)."""

    def __init__(self, gpu_index):
        self.idx, self.stop_flag, self.samples, self.t = gpu_index, threading.Event(), [], None
        self.ok = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, b in bits.items() if any(s[3] & b for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(s[1] for s in self.samples)) if sm else None,
                "power_w_max": max(s[2] for s in self.samples) if sm else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# synthetic workload: layout one-hot, U(-1,1) style images, x_T keyed by global sample index
# ----------------------------------------------------------------------------------------------------------
def synthetic_batch(B, P, n_style, first_index):
    g = torch.Generator().manual_seed(99 + first_index)
    low = torch.rand(B, 1, 8, 8, generator=g)
    mask = (torch.nn.functional.interpolate(low, size=(P, P), mode="bilinear") > 0.5).float()[:, 0]
    seg_oh = torch.stack([1.0 - mask, mask], dim=1)                                  # (B,2,P,P) as the dataloader gives
    style = torch.rand(B, n_style, 3, P, P, generator=torch.Generator().manual_seed(3 + first_index)) * 2 - 1
    L = P // 4
    x_T = torch.stack([torch.randn(3, L, L, generator=torch.Generator().manual_seed(1234 + first_index + i))
                       for i in range(B)])
    img = torch.zeros(B, 3, P, P)
    return img, seg_oh, style, x_T


def build_model(latent, n_style, precision):
    from stedm_b200.config import load_config
    from stedm_b200.modules.ldm_diffusion import LDM_Diffusion
    from stedm_b200.utils.fixture import apply_fixture_weights
    sampling = "mp" if n_style > 1 else "augmented"
    cfg = load_config(["style_agg=mean", f"style_sampling={sampling}", f"ddim_steps={DDIM_STEPS}", f"eta={ETA}",
                       f"cfg_scale={CFG_SCALE}", f"diffusion.image_size={latent}", f"data.patch_size={4 * latent}"])
    if n_style > 1:
        cfg.style_sampling.num_patches = n_style
    m = LDM_Diffusion(cfg, precision=precision, load_first_stage_ckpt=False)
    apply_fixture_weights(m._model, seed=0)          # random-init weights of the architecture (no checkpoints offline)
    return m


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores, bounded sample
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_sample(latent, n_style, sample_batch=1, sample_steps=None, sd=None):
    """One pass of the reference's algorithm (oracle port) for `sample_batch` images on all host threads: conditioning x2,
    guided DDIM steps (all DDIM_STEPS of them when ``sample_steps`` is None — then images/s is plain measured wall time —
    otherwise the first `sample_steps`, with the loop time scaled to DDIM_STEPS and the result marked as extrapolated),
    VQ decode, uint8."""
    from oracle import stedm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    P = 4 * latent
    if sd is None:
        from tests.util import oracle_state_dict
        sd = oracle_state_dict(build_model(latent, n_style, "bf16")._model)
    seg, style, x_T = O.synthetic_batch(sample_batch, P, n_style, seed=11)
    n_steps = DDIM_STEPS if sample_steps is None else sample_steps
    with torch.no_grad():
        t0 = time.perf_counter()
        cond = O.get_conditioning(sd, seg, style)
        unc = O.get_conditioning(sd, seg, torch.zeros_like(style) - 2)
        t1 = time.perf_counter()
        z, _ = O.ddim_sample(sd, cond, unc, x_T, S=DDIM_STEPS, cfg_scale=CFG_SCALE, max_steps=n_steps)
        t2 = time.perf_counter()
        img = O.decode_first_stage(sd, z)
        O.to_uint8(img)
        t3 = time.perf_counter()
    t_cond, t_loop, t_dec = t1 - t0, t2 - t1, t3 - t2
    per_batch = t_cond + t_loop * (DDIM_STEPS / n_steps) + t_dec
    full = n_steps == DDIM_STEPS
    what = (f"all {DDIM_STEPS} guided DDIM steps (2 U-Net passes each), measured wall time" if full else
            f"{n_steps} of {DDIM_STEPS} guided DDIM steps (2 U-Net passes each), loop time scaled x{DDIM_STEPS}/{n_steps}")
    return dict(value=sample_batch / per_batch, seconds_measured=t3 - t0, seconds_per_pass=per_batch, t_cond=t_cond,
                t_ddim_loop=t_loop, t_decode=t_dec, cores=torch.get_num_threads(), extrapolated=not full,
                sample=f"batch {sample_batch} at {P}x{P}: conditioning x2 + {what} + VQ decode")


def run_reference_arm(args, emit):
    """`--impl reference`: the reference's own algorithm (oracle port, fp32, every host thread) on the native arm's
    workload, each step = ONE FULL pass (conditioning x2 + DDIM-50 with guidance + decode) for a batch of `--ref-batch`
    images, timed wall clock — ms_per_step x steps is what the run actually took."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.util import oracle_state_dict
    sd = oracle_state_dict(build_model(args.latent, args.n_style, "bf16")._model)
    secs, last = [], None
    for i in range(args.warmup + args.steps):
        # untimed warm-up passes (thread pool, allocator, oneDNN primitive caches) run 2 of the 50 DDIM steps; every
        # TIMED pass is the full loop
        last = cpu_reference_sample(args.latent, args.n_style, args.ref_batch, 2 if i < args.warmup else None, sd)
        if i >= args.warmup:
            secs.append(last["seconds_per_pass"])
    ms = 1000.0 * float(np.mean(secs))
    v = args.ref_batch / (ms / 1000.0)
    cfg = workload_config(args)
    cfg["batch_per_gpu"] = args.ref_batch      # what this arm ran: the CPU path at the native arm's batch would take hours
    cfg["workload"] = cfg["workload"].replace(f"batch {args.batch} per GPU", f"batch {args.ref_batch} on the host CPU")
    cfg["parallelism"] = f"{last['cores']} host threads, no GPU"
    line = {"metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args):
    P = 4 * args.latent
    return {"workload": f"flowers DDIM-{DDIM_STEPS} cfg {CFG_SCALE} sampling, batch {args.batch} per GPU, {P}x{P} image "
                        f"(latent {args.latent}), style_agg=mean, {args.n_style} style image(s) per sample, "
                        f"random-init fixture weights (BASELINE.json configs[1])",
            "batch_per_gpu": args.batch, "image": P, "ddim_steps": DDIM_STEPS, "cfg_scale": CFG_SCALE, "eta": ETA,
            "parallelism": f"dp{args.gpus} (one generation shard per rank, no per-step collective)",
            "l2": "no flush: every step streams several GB of activations and 0.6 GB of weights, far above the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--n-style", type=int, default=1, dest="n_style")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-graph", action="store_true", help="launch the U-Net pass eagerly instead of replaying the "
                                                            "CUDA graph cached per shape (measured: +2 %% at B=64)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-secondary", action="store_true", help="no HER2 / 512^2 / torch-eager side measurements")
    ap.add_argument("--ref-batch", type=int, default=1, dest="ref_batch",
                    help="images per pass of the CPU reference arm (a full DDIM-50 pass per step)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        return run_reference_arm(args, emit)

    from stedm_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, L, P = args.batch, args.latent, 4 * args.latent

    torch.set_float32_matmul_precision("high")                    # as predict_diff.py:68 (no library GEMM is on the path)
    m = build_model(L, args.n_style, args.precision).to(dev).eval()
    m._model.use_cuda_graph = not args.no_graph                   # captured once per shape, replayed every step
    first = rank * B                                              # global sample index of this rank's shard
    img, seg_oh, style, x_T = synthetic_batch(B, P, args.n_style, first)
    host = [t.pin_memory() for t in (img, seg_oh, style, x_T)]
    devb = [t.to(dev) for t in host]
    gathered = [torch.empty((B, P, P, 3), dtype=torch.uint8, device=dev) for _ in range(world)] if world > 1 else None
    out_host = torch.empty((B, P, P, 3), dtype=torch.uint8).pin_memory()

    def one_pass(bimg, bseg, bstyle, bx):
        batch = (bimg, bseg.clone(), None, bstyle, None)           # prepare_batch edits seg_oh in place
        u8 = m.generate(m.prepare_batch(batch), x_T=bx)
        if world > 1:
            dist.all_gather(gathered, u8)                          # final image gather over NVLink (north_star)
        return u8

    def resident_step():
        return one_pass(*devb)

    def e2e_step():
        d = [t.to(dev, non_blocking=True) for t in host]
        u8 = one_pass(*d)
        out_host.copy_(u8, non_blocking=True)
        return u8

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(k):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for _ in range(args.warmup):
            resident_step()
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        l0 = ops.LAUNCHES[0]
        ms = timed(resident_step, args.steps)
        launches = ops.LAUNCHES[0] - l0
        clk = clocks.stop() if rank == 0 else None
        u8_last = e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
        # rank 0's shard = global samples 0..B-1 with the same seeds at every N: the digest must not depend on N
        digest = hashlib.sha256(u8_last.cpu().numpy().tobytes()).hexdigest() if rank == 0 else None
        strong = None
        if world > 1 and B % world == 0:
            # strong scaling: ONE global batch of B images split over the ranks (rank r takes samples r*B/N ...)
            bs = B // world
            sh = synthetic_batch(bs, P, args.n_style, rank * bs)
            sdev = [t.to(dev) for t in sh]
            part = [torch.empty((bs, P, P, 3), dtype=torch.uint8, device=dev) for _ in range(world)]

            def strong_step():
                u8 = m.generate(m.prepare_batch((sdev[0], sdev[1].clone(), None, sdev[2], None)), x_T=sdev[3])
                dist.all_gather(part, u8)
                return u8

            for _ in range(max(3, args.warmup)):
                strong_step()
            ms_s = timed(strong_step, args.steps)
            strong = {"global_batch": B, "batch_per_gpu": bs, "value": B * args.steps / (ms_s / 1000.0), "unit": "images/s",
                      "ms_per_step": ms_s / args.steps}
        roof = kernel_roofline(m, devb, B, L, args) if rank == 0 else None
        secondary = None
        if rank == 0 and world == 1 and not args.skip_secondary and (B, L, args.n_style) == (64, 64, 1):
            del m
            torch.cuda.empty_cache()
            secondary = secondary_measurements(dev, args)

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    n_img = world * B * args.steps
    value = n_img / (ms / 1000.0)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = out_host.numel()
    line = {"metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": n_img / (ms_e2e / 1000.0), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "unet_step_ms": roof.pop("unet_step_ms"), "unet_step_frac_of_peak": roof.pop("unet_step_frac"),
            "roofline": roof, "images_sha256": digest}
    if strong is not None:
        line["strong_scaling"] = strong
    if secondary is not None:
        line["secondary"] = secondary
    if world == 1 and not args.skip_cpu_baseline:
        cb = cpu_reference_sample(L, args.n_style, 1, 2)      # bounded: 2 guided steps of 50, scaled (the reference arm
        # of this script, --impl reference, times full DDIM-50 passes)
        line["cpu_baseline"] = {"value": cb["value"], "unit": "images/s", "cores": cb["cores"], "kind": "port",
                                "sample": cb["sample"]}
    emit(line)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the dominant kernel's 1024->1024 3x3 @16x16
    launch on 128 samples, read from the newest committed `ncu --set full` summary profiles/rNN_ncu_full_conv_tc.csv
    (first kernel row; algorithmic bytes of that launch: 153 MB = 67 MB input + 19 MB weights + 67 MB output)."""
    import csv
    import glob
    import re
    files = [f for f in glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_conv_tc.csv")) if re.search(r"r\d+_ncu_full_conv_tc\.csv$", f)]
    if not files:
        return None
    try:
        with open(sorted(files)[-1], newline="") as fh:
            rows = list(csv.reader(fh))
        head, units, first = rows[0], rows[1], rows[2]
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = head.index(name)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            tot += float(first[i]) * scale
        return tot
    except Exception:
        return None


def secondary_measurements(dev, args):
    """Driver-visible side numbers (rank 0, N = 1 only; each a short run after the headline measurement):
    BASELINE configs[2] (HER2-style N = 10 style images per sample), configs[3] geometry on one GPU (512^2, latent 128),
    and configs[4]'s kernel bar: the U-Net eps step of the reference's algorithm under torch eager (cuDNN / cuBLAS, TF32
    as predict_diff.py:68 sets, and bf16 autocast) on this same B200 next to the native pass."""
    out = {}

    def run_cfg(latent, n_style, batch, steps):
        P = 4 * latent
        mm = build_model(latent, n_style, args.precision).to(dev).eval()
        mm._model.use_cuda_graph = not args.no_graph
        b = [t.to(dev) for t in synthetic_batch(batch, P, n_style, 0)]
        fn = lambda: mm.generate(mm.prepare_batch((b[0], b[1].clone(), None, b[2], None)), x_T=b[3])
        with torch.no_grad():
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return mm, {"value": batch / (ms / 1000.0), "unit": "images/s", "ms_per_step": ms, "batch": batch, "image": P,
                    "style_images_per_sample": n_style, "steps": steps, "warmup": 3}

    mm, out["her2_n10_256"] = run_cfg(64, 10, 64, 2)
    # configs[4]: one U-Net eps pass, native (graph replay, as the sampler runs it) vs torch eager on the same GPU
    from oracle import stedm_oracle as O
    from tests.util import oracle_state_dict
    unet = mm._model.model.diffusion_model
    sd = {k: v.to(dev) for k, v in oracle_state_dict(mm._model).items() if k.startswith(O.UNET)}
    g = torch.Generator().manual_seed(5)
    Bq = 64
    x, cc = torch.randn(Bq, 3, 64, 64, generator=g).to(dev), torch.randn(Bq, 3, 64, 64, generator=g).to(dev)
    ctx, t = torch.randn(Bq, 512, generator=g).to(dev), torch.full((Bq,), 481, dtype=torch.long, device=dev)

    def time_ms(fn, warm, reps):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.no_grad():
        native = time_ms(lambda: unet.forward_split(x, cc, t, ctx), 5, 10)
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        xc = torch.cat([x, cc], 1)
        eager32 = time_ms(lambda: O.unet_forward(sd, xc, t, ctx), 12, 10)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            eager16 = time_ms(lambda: O.unet_forward(sd, xc, t, ctx), 12, 10)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = tf32
    out["torch_eager_b200"] = {"what": "one U-Net eps pass, batch 64, latent 64 (217.29 GFLOP per sample): the reference's "
                                       "UNetModel.forward restated functionally (oracle, pinned bit-equal to the reference on "
                                       "CPU) under torch eager on this GPU, 12 warm-up passes, vs the native pass",
                               "native_ms": native, "torch_tf32_ms": eager32, "torch_bf16_autocast_ms": eager16,
                               "speedup_vs_tf32": eager32 / native, "speedup_vs_bf16_autocast": eager16 / native}
    del mm, unet, sd
    torch.cuda.empty_cache()
    mm, out["catch_512"] = run_cfg(128, 1, 64, 1)
    del mm
    torch.cuda.empty_cache()
    return out


def kernel_roofline(m, devb, B, L, args):
    """CUDA-event timing of every tensor-core convolution launch of one guided DDIM step (eager, after the timed
    region, same stream) -> achieved TFLOP/s of the dominant kernel (stedm_conv_tc), plus the U-Net step time."""
    from stedm_b200 import ops
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    model = m._model
    pk = peaks()
    img, seg_oh, style, x_T = devb
    batch = m.prepare_batch((img, seg_oh.clone(), None, style, None))
    _, c = model.get_input(batch, "image")
    _, cu = model.get_input(dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2), "image")
    sampler = DDIMSampler(model, use_cuda_graph=False)
    sampler.make_schedule(ddim_num_steps=DDIM_STEPS, ddim_eta=ETA, verbose=False)
    ts = torch.full((B,), 481, device=x_T.device, dtype=torch.long)
    # whole-step time (2B-sample U-Net pass + K11), eager, averaged
    for _ in range(2):
        sampler.p_sample_ddim(x_T, c, ts, index=24, unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=cu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        sampler.p_sample_ddim(x_T, c, ts, index=24, unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=cu)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 3
    step_flops = 2 * B * UNET_GFLOP_PER_SAMPLE_L64 * (L / 64.0) ** 2 * 1e9
    # per-launch timing of the tensor-core convolution
    rec = []
    orig = ops.conv

    def timed_conv(x0, weight, bias, cout, ksize, **kw):
        if not kw.get("tensor_core", True):
            return orig(x0, weight, bias, cout, ksize, **kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = orig(x0, weight, bias, cout, ksize, **kw)
        b.record()
        bsz, h, w, c0 = x0.shape
        c1 = 0 if kw.get("x1") is None else kw["x1"].shape[-1]
        taps = 4 if kw.get("up_phase") is not None else ksize * ksize   # executed taps (folded upsample: 4, not 9)
        sk = kw.get("skip_x0")      # fused 1x1 skip convolution: extra K columns
        skc = 0 if sk is None else sk.shape[-1] + (0 if kw.get("skip_x1") is None else kw["skip_x1"].shape[-1])
        st = kw.get("stride", 1)                                        # stride 2: GEMM rows = output pixels
        rec.append((a, b, 2.0 * bsz * (h // st) * (w // st) * cout * (taps * (c0 + c1) + skc), cout))
        return out

    ops.conv = timed_conv
    try:
        sampler.p_sample_ddim(x_T, c, ts, index=24, unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=cu)
        torch.cuda.synchronize()
    finally:
        ops.conv = orig
    tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in rec)
    tot_fl = sum(f for _, _, f, _ in rec)
    big = [(a.elapsed_time(b), f) for a, b, f, co in rec if co % 256 == 0]
    big_ms, big_fl = sum(t for t, _ in big), sum(f for _, f in big)
    achieved = big_fl / (big_ms / 1e3) / 1e12 if big_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "conv_tc_kernel<256> (tcgen05 implicit-GEMM conv, BN=256)",
            "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']}): cuBLAS bf16 8192^3 back to back for "
                           f"4 s under the same power cap; the kernel is timed inside the long step, so this is the "
                           f"applicable denominator (frac > 1 = faster than that cuBLAS run)",
            "frac_of_burst_peak": achieved / pk["tf_burst"], "burst_peak": pk["tf_burst"],
            # dram__bytes_read.sum + dram__bytes_write.sum of the 1024->1024 3x3 @16x16 (128 samples) launch from the
            # committed `ncu --set full` capture profiles/r01_ncu_full_conv_tc.csv, row 1 (algorithmic bytes: 153 MB =
            # 67 MB input + 19 MB weights + 67 MB output; part of the input is still L2-resident from its producer)
            "traffic": ncu_traffic_bytes() if (B == 64 and L == 64) else None,
            "launches_timed": len(big), "avg_launch_ms": big_ms / max(1, len(big)),
            "algorithmic_flops_per_launch": big_fl / max(1, len(big)),
            "all_tc_conv": {"launches": len(rec), "ms": tot_ms, "tflops": tot_fl / (tot_ms / 1e3) / 1e12 if tot_ms else 0.0,
                            "share_of_unet_step": tot_ms / step_ms},
            # executed work only: FLOPs skipped by the shared encoder trunk are not credited
            "executed_conv_tflop_per_step": tot_fl / 1e12, "unshared_tflop_per_step": step_flops / 1e12,
            "unet_step_ms": step_ms, "unet_step_frac": tot_fl / (step_ms / 1e3) / 1e12 / pk["tf_sustained"]}


if __name__ == "__main__":
    main()
