"""Is anything hidden between the phases of a generation pass?  CUDA-event / wall time of LDM_Diffusion.generate() next to the
synchronised phase times (conditioning x2, DDIM-50 loop, decode) and the replay of the captured loop graph alone, on one box.
    python tools/gap_check.py   (round 2: generate() 938 ms = phases 937 ms; loop graph 890 ms = 50 x 17.8 ms)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
B, L = 64, 64
m = bench.build_model(L, 1, "bf16").to(dev).eval()
model = m._model
img, seg, style, x_T = [t.to(dev) for t in bench.synthetic_batch(B, 4 * L, 1, 0)]
def full():
    return m.generate(m.prepare_batch((img, seg.clone(), None, style, None)), x_T=x_T)
with torch.no_grad():
    for _ in range(4): full()
    torch.cuda.synchronize()
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); full(); e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"generate(): events {e0.elapsed_time(e1):.1f} ms, wall {1e3*(t1-t0):.1f} ms")
    # phases
    def st(fn):
        torch.cuda.synchronize(); t = time.perf_counter(); o = fn(); torch.cuda.synchronize(); return o, 1e3*(time.perf_counter()-t)
    for it in range(2):
        batch = m.prepare_batch((img, seg.clone(), None, style, None))
        (z, c), tc = st(lambda: model.get_input(batch, "image"))
        unc = dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2)
        (_, cu), tu = st(lambda: model.get_input(unc, "image"))
        (out, _), ts = st(lambda: model.sample_log(c, batch_size=B, ddim=True, ddim_steps=50, eta=0.0, log_every_t=1000, x_T=x_T, unconditional_conditioning=cu, unconditional_guidance_scale=1.5))
        dec, td = st(lambda: model.decode_first_stage(out))
        print(f"phases: cond {tc:.1f} uncond {tu:.1f} ddim50 {ts:.1f} decode {td:.1f} sum {tc+tu+ts+td:.1f}")
    # the loop graph alone
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    unet = model.model.diffusion_model
    cache = unet.__dict__.get("_graph_cache", {})
    for k, ent in cache.items():
        if isinstance(k, tuple) and k and k[0] == "loop":
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ent["graph"].replay(); e1.record(); torch.cuda.synchronize()
            print(f"loop graph replay alone: {e0.elapsed_time(e1):.1f} ms ({ent['launches']} launches)")
