"""Wall-clock (synchronised) time of each phase of one generation pass at the bench workload."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def sync_time(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) * 1e3


def main():
    B, L = int(os.environ.get("B", 64)), 64
    dev = torch.device("cuda", 0)
    m = bench.build_model(L, 1, "bf16").to(dev).eval()
    m._model.use_cuda_graph = os.environ.get("GRAPH", "1") == "1"
    model = m._model
    host = [t.pin_memory() for t in bench.synthetic_batch(B, 4 * L, 1, 0)]
    with torch.no_grad():
        for it in range(4):
            (img, seg, style, x_T), t_h2d = sync_time(lambda: [t.to(dev, non_blocking=True) for t in host])
            batch = m.prepare_batch((img, seg.clone(), None, style, None))
            (z, c), t_c = sync_time(lambda: model.get_input(batch, "image"))
            unc = dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2)
            (_, cu), t_u = sync_time(lambda: model.get_input(unc, "image"))
            (out, _), t_s = sync_time(lambda: model.sample_log(c, batch_size=B, ddim=True, ddim_steps=50, eta=0.0,
                                                               log_every_t=1000, x_T=x_T, unconditional_conditioning=cu,
                                                               unconditional_guidance_scale=1.5))
            dec, t_d = sync_time(lambda: model.decode_first_stage(out))
            from stedm_b200 import ops
            u8, t_8 = sync_time(lambda: ops.image_to_uint8(dec.contiguous()))
            _, t_d2h = sync_time(lambda: u8.cpu())
            print(f"iter {it}: h2d {t_h2d:.1f}  cond {t_c:.1f}  uncond {t_u:.1f}  ddim50 {t_s:.1f}  decode {t_d:.1f}  "
                  f"u8 {t_8:.2f}  d2h {t_d2h:.1f} ms;  mem alloc {torch.cuda.memory_allocated() / 2**30:.1f} GiB "
                  f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)


if __name__ == "__main__":
    main()
