#!/usr/bin/env python
"""Host<->device copy ceiling of the box with ALL ranks copying at once and NO kernels running — the denominator for
bench.py's end-to-end scaling: frames/s at N GPUs cannot exceed
N x (H2D GB/s per rank) / bytes per frame.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      tools/h2d_ceiling.py [--mb 201] [--seconds 0.6]

Variants per rank (each timed with CUDA events on its own streams, all ranks inside one barrier pair):
  buffer mode   pinned | wc | huge       (drnb200.frameio.HostBuffer, include/drnb200.h)
  chunks        1 | 4                    (the batch split into row chunks on separate streams)
  direction     h2d | h2d+d2h            (label maps going back at 1/24 (fp32 frames) of the input bytes)
Rank 0 prints one JSON line per variant: per-rank min/mean/max GB/s and the aggregate.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-seg-model-compress_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from drnb200.frameio import HostBuffer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=201.3, help="bytes per H2D batch (201.3 = 8 fp32 1024x2048 frames)")
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--modes", default="pinned,wc,huge")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        os.dup2(saved, 1)
    nbytes = int(args.mb * 1e6) // 4096 * 4096
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes // 24, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    back = torch.cuda.Stream(device=dev)
    h_out = HostBuffer((nbytes // 24,), torch.uint8, "pinned")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for mode in args.modes.split(","):
        try:
            t0 = time.perf_counter()
            hb = HostBuffer((nbytes,), torch.uint8, mode)
            hb.tensor[::4096] = 1                       # touch every page from THIS process
            alloc_s = time.perf_counter() - t0
        except Exception as exc:                        # mode not available on this host
            if rank == 0:
                print(json.dumps({"mode": mode, "error": str(exc)[:200]}))
            continue
        for chunks in (1, 4):
            for bidir in (False, True):
                step = nbytes // chunks

                def one():
                    for c in range(chunks):
                        with torch.cuda.stream(streams[c]):
                            d_in[c * step:(c + 1) * step].copy_(hb.tensor[c * step:(c + 1) * step], non_blocking=True)
                    if bidir:
                        with torch.cuda.stream(back):
                            h_out.tensor.copy_(d_out, non_blocking=True)

                for _ in range(2):
                    one()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                # size the loop for ~args.seconds
                e0.record()
                one()
                for s in streams[:chunks] + ([back] if bidir else []):
                    torch.cuda.current_stream().wait_stream(s)
                e1.record()
                torch.cuda.synchronize()
                reps = max(3, int(args.seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
                barrier()
                e0.record()
                for _ in range(reps):
                    one()
                for s in streams[:chunks] + ([back] if bidir else []):
                    torch.cuda.current_stream().wait_stream(s)
                e1.record()
                barrier()
                gbs = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
                t = torch.tensor([gbs], device=dev, dtype=torch.float64)
                if world > 1:
                    all_t = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(all_t, t)
                    vals = [float(v) for v in all_t]
                else:
                    vals = [gbs]
                if rank == 0:
                    print(json.dumps({"mode": mode, "chunks": chunks, "d2h_too": bidir, "n_gpus": world,
                                      "h2d_gbs_per_rank": [round(v, 2) for v in vals],
                                      "h2d_gbs_sum": round(sum(vals), 1), "h2d_gbs_min": round(min(vals), 2),
                                      "batch_mb": nbytes / 1e6, "reps": reps, "alloc_s": round(alloc_s, 3)}), flush=True)
        hb.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
