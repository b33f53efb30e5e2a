#!/usr/bin/env python
"""What limits the end-to-end (host frames -> labels on the host) rate at N GPUs?  Variants of bench.py's e2e pipeline on
uint8 frames, all ranks at once, max over ranks:
  base     2 input buffers, labels D2H on their own stream (what bench.py does)
  no_d2h   no label read-back at all (H2D + kernels only)
  hist     evaluation mode: the confusion matrix is updated on the device, 2.9 kB read back per step
  buf3     3 input buffers
  serial   labels D2H queued on the H2D stream (the two directions never overlap)
  python -m torch.distributed.run --nproc-per-node N ... tools/e2e_variants.py [--steps 20]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-seg-model-compress_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--f32", action="store_true", help="float32 NCHW frames instead of uint8 HWC")
    a = ap.parse_args()
    args = bench.parse([])
    import drnb200
    from drnb200 import synthetic
    ctx = bench.Ctx()
    dev = ctx.dev
    model, _, _ = bench.build_model(args, dev)
    B, H, W = args.batch, args.height, args.width
    model.set_ingest(bench.INFO_MEAN, bench.INFO_STD)
    if a.f32:
        hx = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
        hx.normal_()
    else:
        hx = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        hx.copy_(synthetic.make_u8_frames(B, H, W, seed=1234 + ctx.rank))
    gt = torch.randint(0, 19, (B, H, W), device=dev, dtype=torch.int64).to(torch.uint8)
    meter = drnb200.ConfusionMeter(19, dev)
    main_stream = torch.cuda.current_stream()

    def run(variant, steps):
        nbuf = 3 if variant == "buf3" else 2
        copy_stream = torch.cuda.Stream(device=dev)
        d2h_stream = copy_stream if variant == "serial" else torch.cuda.Stream(device=dev)
        hl = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        hh = [torch.empty((19, 19), dtype=torch.int64).pin_memory() for _ in range(nbuf)]
        xd = [torch.empty(hx.shape, dtype=hx.dtype, device=dev) for _ in range(nbuf)]
        ready = [torch.cuda.Event() for _ in range(nbuf)]
        consumed = [torch.cuda.Event() for _ in range(nbuf)]
        d2h_done = [torch.cuda.Event() for _ in range(nbuf)]
        keep = [None] * nbuf
        for b in range(nbuf):
            consumed[b].record(main_stream)
            d2h_done[b].record(main_stream)

        def loop(n):
            for s_ in range(n + 1):
                if s_ < n:
                    b = s_ % nbuf
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(consumed[b])
                        xd[b].copy_(hx, non_blocking=True)
                        ready[b].record(copy_stream)
                if s_ >= 1:
                    b = (s_ - 1) % nbuf
                    main_stream.wait_event(ready[b])
                    main_stream.wait_event(d2h_done[b])
                    keep[b] = labels = model.predict(xd[b])
                    if variant == "hist":
                        meter.update(labels, gt)
                    consumed[b].record(main_stream)
                    if variant == "no_d2h":
                        continue
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(consumed[b])
                        if variant == "hist":
                            hh[b].copy_(meter.hist, non_blocking=True)
                        else:
                            hl[b].copy_(labels, non_blocking=True)
                        d2h_done[b].record(d2h_stream)

        loop(3)
        ctx.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        loop(steps)
        main_stream.wait_stream(d2h_stream)
        f1.record()
        ctx.barrier()
        wall = time.perf_counter() - t0
        return max(f0.elapsed_time(f1), 1e3 * wall)

    with torch.no_grad():
        for _ in range(3):
            model.predict(hx.to(dev))
        out = {}
        for variant in ("base", "no_d2h", "hist", "buf3", "serial", "base"):
            ms = ctx.max_over_ranks([run(variant, a.steps)])[0]
            per_rank = ctx.gather(ms)
            key = variant if variant not in out else variant + "_again"
            out[key] = {"frames_per_s": ctx.world * B * a.steps / (ms * 1e-3), "ms_per_step": ms / a.steps}
    if ctx.rank == 0:
        print(json.dumps({"n_gpus": ctx.world, "frames": "f32" if a.f32 else "uint8", "steps": a.steps, "variants": out}))
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
