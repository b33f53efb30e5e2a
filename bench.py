#!/usr/bin/env python
"""bench.py — headline benchmark: block-pruned DRN-D-22 frames/s @1024x2048 on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference path on the host CPU cores)

A step is one pass of the hot path over one batch of synthetic frames per GPU (BASELINE.json configs[1]:
DRN-D-22, BlockPruner 75 % block-sparse masks, batch 8, 1024x2048).  Frames shard across ranks with no
data-path collective (weights/tile lists replicated, "weak" scaling); the only NCCL call is the all-reduce of
the 19x19 int64 confusion matrix at the end of the evaluation, inside the timed region.
One JSON line is printed by rank 0; see README/DESIGN.md for the keys (roofline, cpu_baseline, e2e, clocks).
"""
import argparse
import collections
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "video-seg-model-compress_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "pruned DRN-D-22 frames/s @1024x2048"      # BASELINE.json metric (default arguments)


def metric_name(args):
    """the BASELINE metric for the default workload; other --arch/--height/--width runs are labelled as what they are"""
    if args.arch == "drn_d_22" and (args.height, args.width) == (1024, 2048):
        return METRIC
    return "pruned %s frames/s @%dx%d" % (args.arch.upper().replace("_", "-"), args.height, args.width)
# per-channel statistics of the reference's video set (info.json of the reference)
INFO_MEAN = (0.29010095242892997, 0.32808144844279574, 0.28696394422942517)
INFO_STD = (0.1829540508368939, 0.18656561047509476, 0.18447508988480435)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="drn_d_22")
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--sparsity", type=float, default=0.75)
    ap.add_argument("--act", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit activation storage; fp16 is what passes the 99.9 %% label gate (DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers-out", default=None, help="write the per-layer timing table (JSON) here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "src": "fallback"}


def build_model(args, device=None):
    """synthetic weights of the named architecture + BlockPruner masks; returns (model, sd, pruner)"""
    import drnb200
    from drnb200 import synthetic
    model = drnb200.DRNSeg(args.arch, 19, pretrained_model=None, pretrained=False, act_dtype=args.act)
    shapes = collections.OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    sd = synthetic.make_state_dict(shapes, seed=0)
    model.load_state_dict(sd, strict=False)
    pruner = None
    if args.sparsity > 0:
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
            json.dump(synthetic.block_pruner_config(shapes, args.sparsity), fh)
        pruner = drnb200.pruners.make_pruner(fh.name, on_gpu=False)
        pruner.generate_masks(model, is_static=False)          # the reference's magnitude block pruning
        os.unlink(fh.name)
        sd = synthetic.sparse_reinit(sd, pruner.mask_dict, seed=0)
        model.load_state_dict(sd, strict=False)
        pruner.apply_masks(model)
    if device is not None:
        model = model.to(device).eval()
        model.set_pruner(pruner)
    return model, sd, pruner


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                 "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, b in names.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args):
    """the reference's CPU implementation of the path (oracle port: torch CPU fp32, all host threads), timed on
    a bounded sample: one 1024x2048 frame per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import drn_oracle
    _, sd, _ = build_model(args)
    from drnb200 import synthetic
    x = synthetic.make_frames(1, args.height, args.width, seed=1234)
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone and takes the whole host
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    threads = torch.get_num_threads()
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            torch.max(drn_oracle.drnseg_forward(sd, x)[0], 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            final, _ = drn_oracle.drnseg_forward(sd, x)
            _, pred = torch.max(final, 1)           # semantic_seg.py:444-445
        dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = "1 frame %dx%d per step, fp32, %d steps" % (args.height, args.width, args.steps)
    line = {"impl": "reference", "metric": metric_name(args), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload(args, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pin_to_gpu_numa_node(local_rank):
    """multi-rank runs: bind this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host
    buffers are allocated (first touch puts their pages on that node), so every rank's H2D stream reads local DRAM.
    Returns a short description for the JSON line; does nothing when the topology is not exposed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                       # nvml pads the PCI domain to 8 hex digits
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return "numa_node not exposed"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "node %d has no CPU in this process's affinity mask" % node
        os.sched_setaffinity(0, cpus)
        return "rank bound to NUMA node %d (%d CPUs)" % (node, len(cpus))
    except Exception as exc:                                  # topology files missing, no NVML, no permission
        return "not bound (%s)" % type(exc).__name__


def workload(args, batch):
    return {"workload": "%s BlockPruner %.0f%% block-sparse, batch %d/GPU, %dx%d, act %s" % (
        args.arch, 100 * args.sparsity, batch, args.height, args.width, args.act),
        "arch": args.arch, "batch_per_gpu": batch, "height": args.height, "width": args.width,
        "sparsity": args.sparsity, "act_dtype": args.act,
        "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2; no explicit flush" % (
            batch * 3 * args.height * args.width * 4 / 1e6)}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import drnb200
    from drnb200 import synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    numa = pin_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        import torch.distributed as dist
        # NCCL printf()s its version banner to stdout while the communicator is created: keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    model, sd, pruner = build_model(args, dev)
    B, H, W = args.batch, args.height, args.width
    # synthetic frames: a distinct batch per rank, generated on the device outside the timed region
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, 3, H, W, device=dev, generator=g)
    gt = torch.randint(0, 19, (B, H, W), device=dev, generator=g, dtype=torch.int64).to(torch.uint8)
    meter = drnb200.ConfusionMeter(19, dev)
    eng = model.engine()

    def step(inp):
        labels = model.predict(inp)
        meter.update(labels, gt)
        return labels

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step(x)
        if dist is not None:
            meter.all_reduce()          # warm NCCL
        barrier()
        meter.hist.zero_()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step(x)
        meter.all_reduce()              # the evaluation's single collective (NCCL over NVLink)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = (eng.launches_per_forward + 1) * args.steps
        miou = meter.miou()

        # ---- e2e: host frames -> H2D -> predict -> D2H labels, through the public API.
        # A user-side double buffer: frame batch i+1 is copied in on a side stream while batch i is
        # segmented; every step still moves its own input from pinned host memory and its labels back.
        copy_stream = torch.cuda.Stream(device=dev)
        d2h_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()

        def e2e_measure(hx):
            """hx: pinned host batch (float32 NCHW, or uint8 NHWC with set_ingest) -> ms for args.steps steps"""
            hl = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
            xd = [torch.empty(hx.shape, dtype=hx.dtype, device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            d2h_done = [torch.cuda.Event() for _ in range(2)]
            keep = [None, None]

            def e2e_run(n_steps):
                for s_ in range(n_steps + 1):
                    if s_ < n_steps:                      # stage batch s_ on the copy stream
                        b = s_ & 1
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(consumed[b])
                            xd[b].copy_(hx, non_blocking=True)
                            ready[b].record(copy_stream)
                    if s_ >= 1:                           # segment batch s_-1 on the main stream
                        b = (s_ - 1) & 1
                        main_stream.wait_event(ready[b])
                        main_stream.wait_event(d2h_done[b])      # labels of batch s_-3 are on the host:
                        keep[b] = labels = model.predict(xd[b])  # their device buffer may be recycled
                        consumed[b].record(main_stream)
                        with torch.cuda.stream(d2h_stream):      # labels go back on their own stream so the
                            d2h_stream.wait_event(consumed[b])   # next batch's kernels are not queued behind PCIe
                            hl[b].copy_(labels, non_blocking=True)
                            d2h_done[b].record(d2h_stream)

            for b in range(2):
                consumed[b].record(main_stream)
                d2h_done[b].record(main_stream)
            e2e_run(2)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            f0.record()
            e2e_run(args.steps)
            main_stream.wait_stream(d2h_stream)               # the last labels have reached the host
            f1.record()
            barrier()
            wall = time.perf_counter() - t0
            return max(f0.elapsed_time(f1), 1e3 * wall)

        hx = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
        hx.copy_(x.cpu())
        e2e_ms = e2e_measure(hx)
        # the same through the fused frame ingest: uint8 HWC frames as cv2 delivers them, the
        # reference's ToTensor + Normalize (info.json statistics) applied inside the stem kernel
        e2e_u8_ms = None
        if W % 16 == 0:
            model.set_ingest(INFO_MEAN, INFO_STD)
            hu8 = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
            hu8.copy_(synthetic.make_u8_frames(B, H, W, seed=1234 + rank))
            e2e_u8_ms = e2e_measure(hu8)
        sampler.stop_flag = True
        sampler.join()

        # ---- per-layer timing pass (CUDA events per launch) for the roofline objects
        per_layer = collections.OrderedDict()
        reps = 3
        for _ in range(reps):
            tl = []
            eng.run(x, want_labels=True, timings=tl)
            torch.cuda.synchronize()
            for name, a, b in tl:
                per_layer[name] = per_layer.get(name, 0.0) + a.elapsed_time(b) / reps

    # max over ranks
    t = torch.tensor([ms, e2e_ms, e2e_u8_ms or 0.0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_u8_ms = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        pk = peaks()
        frames = world * B * args.steps
        dense, live, tile = eng.mac_counts(B, H, W)
        conv_ms = sum(v for k, v in per_layer.items() if k not in ("stem", "head"))
        conv_live = live - B * H * W * 16 * 147 - B * (H // 8) * (W // 8) * 512 * 19      # minus stem and seg
        conv_tflops = 2.0 * conv_live / (conv_ms * 1e-3) / 1e12
        head_bytes = B * ((H // 8) * (W // 8) * 512 * 2 + H * W) + 19 * 512 * 4 + 19 * 4
        head_gbs = head_bytes / (per_layer["head"] * 1e-3) / 1e9
        layers = []
        lib = drnb200.ffi.lib()
        KERNELS = {0: "conv_tc<T>", 1: "conv_tc<P>", 2: "conv_tc<T,f32>", 3: "conv_gather", 4: "conv_halo",
                   5: "conv_tc<T,ROW>", -1: "conv_direct"}
        shapes = {-1: (H, W)}
        for i, op in enumerate(eng.last_ops or eng.ops):
            src = op.input_from if op.input_from is not None else i - 1
            oh, ow = op.out_hw(*shapes[src])
            shapes[i] = (oh, ow)
            macs = B * oh * ow * op.live_elems
            c = op.conv
            io_bytes = 2 * B * (shapes[src][0] * shapes[src][1] * c.in_channels + oh * ow * op.out_channels)
            lms = per_layer[op.key]
            mode = max([lib.drnb200_conv_plan_mode(pl) for pl in op.plans.values()] or [-1])
            layers.append({"layer": op.key, "ms": lms, "live_gmac": macs / 1e9, "kernel": KERNELS.get(mode, "?"),
                           "tflops_live": 2 * macs / (lms * 1e-3) / 1e12,
                           "tensor_frac": 2 * macs / (lms * 1e-3) / 1e12 / pk["tflops"],
                           "hbm_gbs": io_bytes / (lms * 1e-3) / 1e9,
                           "live_tiles": op.n_live, "tile": [op.tile_o, op.tile_ci]})
        # the dominant kernel = the one with the largest share of the step (conv_tc_kernel<MODE_T, ROW>: the 3x3
        # stride-1 convs of layers 4-8), aggregated over its launches of one step
        by_kernel = collections.defaultdict(float)
        for l in layers:
            by_kernel[l["kernel"]] += l["ms"]
        dom_name = max(by_kernel, key=by_kernel.get)
        dom = [l for l in layers if l["kernel"] == dom_name]
        dom_ms = sum(l["ms"] for l in dom)
        dom_tflops = 2.0 * sum(l["live_gmac"] for l in dom) * 1e9 / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0
        # DRAM traffic per launch from the committed `ncu --set full` capture of the same workload
        # (tools/ncu_summarize.py traffic -> profiles/ncu_traffic.json); null when there is none for this workload
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = {k: v for k, v in json.load(fh).items() if v.get("workload") == workload(args, B)["workload"]}
        dom_traffic = traffic.get("conv_tc_row", {}).get("bytes_per_launch") if "ROW" in dom_name else None
        dom_io = sum(l["hbm_gbs"] * l["ms"] * 1e6 for l in dom) / max(len(dom), 1)    # algorithmic bytes per launch
        if args.layers_out:
            with open(args.layers_out, "w") as fh:
                json.dump({"per_layer_ms": per_layer, "layers": layers, "batch": B}, fh, indent=1)
        line = {
            "metric": metric_name(args), "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.act,
            "data": "synthetic", "config": workload(args, B),
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": B * 3 * H * W * 4, "d2h_bytes_per_step": B * H * W},
            "e2e_uint8_frames": None if not e2e_u8_ms else {
                "value": frames / (e2e_u8_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 3,
                "d2h_bytes_per_step": B * H * W,
                "note": "same pipeline fed with uint8 HWC frames; ToTensor+Normalize fused into the stem kernel"},
            "gpu_launches": launches,
            "host_numa": numa,
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor",
                         "kernel": "%s (%d launches per step: %s)" % (
                             dom_name, len(dom), ", ".join(l["layer"].replace("layer.", "") for l in dom)),
                         "achieved": dom_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": dom_tflops / pk["tflops"], "traffic": dom_traffic,
                         "traffic_note": "avg DRAM bytes read+written per launch (ncu --set full, profiles/ncu_traffic.json)"
                                         "; algorithmic input+output bytes per launch (residual reads excluded): %.0f" % dom_io,
                         "peak_source": pk["src"] + " bf16_tflops_sustained",
                         "flops": "2 x unpruned (mask != 0) MACs of these launches",
                         "ms_per_step": dom_ms},
            "roofline_all_convs": {"bound": "tensor", "kernel": "all %d conv launches of one step (stem/seg excluded)" % len(layers),
                                   "achieved": conv_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                                   "frac": conv_tflops / pk["tflops"], "ms_per_step": conv_ms},
            "roofline_head": {"bound": "hbm", "kernel": "head_fused_kernel (classifier GEMM + x8 upsample + argmax, one launch)",
                              "achieved": head_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": head_gbs / pk["hbm_gbs"],
                              "traffic": traffic.get("head_fused", {}).get("bytes_per_launch"),
                              "bytes_per_launch": head_bytes, "ms": per_layer["head"]},
            "stem_ms_per_step": per_layer["stem"],
            "macs_per_frame": {"dense_g": dense / B / 1e9, "unpruned_g": live / B / 1e9,
                               "live_tile_g": tile / B / 1e9},
            "miou_vs_random_labels": miou,
        }
        if not args.no_cpu_baseline and world == 1:      # reported at N=1 only (rank 0 owns the host cores there)
            from oracle import drn_oracle           # cpu_baseline leg: the oracle port, bounded sample
            xs = synthetic.make_frames(1, H, W, seed=1234)
            with torch.no_grad():
                torch.max(drn_oracle.drnseg_forward(sd, xs)[0], 1)
                t0 = time.perf_counter()
                n_it = 3
                for _ in range(n_it):
                    torch.max(drn_oracle.drnseg_forward(sd, xs)[0], 1)
                dt = (time.perf_counter() - t0) / n_it
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": "1 frame %dx%d x %d iterations, torch CPU fp32 (oracle port of "
                                              "semantic_seg.DRNSeg + torch.max)" % (H, W, n_it)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
