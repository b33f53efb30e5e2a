#!/usr/bin/env python
"""bench.py — headline benchmark: block-pruned DRN-D-22 frames/s @1024x2048 on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference path on the host CPU cores)

A step is one pass of the hot path over one batch of synthetic frames per GPU (BASELINE.json configs[1]:
DRN-D-22, BlockPruner 75 % block-sparse masks, batch 8, 1024x2048).  Frames shard across ranks with no
data-path collective (weights/tile lists replicated, "weak" scaling); the only NCCL call is the all-reduce of
the 19x19 int64 confusion matrix at the end of the evaluation, inside the timed region.
One JSON line is printed by rank 0.  Beside the contract keys it carries
  roofline / roofline_all_convs / roofline_head   fractions of the BURST peaks for event-timed kernels (+ of sustained)
  e2e / e2e_uint8_frames     drnb200.FramePipeline: pinned host frames -> H2D -> predict -> labels D2H every step, with
                             the per-rank copy rates
  host_ceiling               H2D GB/s with ALL ranks copying and no kernels: the bound of e2e scaling on this host
  bf16                       the same workload with bf16 activation storage (frames/s + label agreement)
  sustained                  >= 3 s of back-to-back steps: frames/s, SM clock, power, throttle reasons
  configs                    short runs of BASELINE configs 3/4/5 on their own masks (--pruner rmb/srmbrep/unstructured)
  cpu_baseline / parity      the oracle port on the host cores (N=1 only) and the label agreement against it
"""
import argparse
import collections
import contextlib
import io
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
# per-channel statistics of the reference's video set (info.json of the reference)
INFO_MEAN = (0.29010095242892997, 0.32808144844279574, 0.28696394422942517)
INFO_STD = (0.1829540508368939, 0.18656561047509476, 0.18447508988480435)
PRUNER_LABEL = {"block": "BlockPruner %.0f%% block-sparse", "rmb": "RmbPruner outer 50%% + blocklets",
                "srmbrep": "srmbrep %.0f%% inner (optimal_configs shape)", "unstructured": "l1_unstructured 90%%",
                "none": "dense"}


def metric_name(args):
    """the BASELINE metric for the default workload; other runs are labelled as what they are"""
    if args.arch == "drn_d_22" and (args.height, args.width) == (1024, 2048) and args.pruner == "block":
        return METRIC
    return "pruned %s (%s) frames/s @%dx%d" % (args.arch.upper().replace("_", "-"), args.pruner, args.height, args.width)


def parse(argv=None):
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
    ap.add_argument("--pruner", default="block", choices=["block", "rmb", "srmbrep", "unstructured", "none"],
                    help="mask source: BlockPruner (config 2), RmbPruner (config 3), srmbrep (config 4), "
                         "torch l1_unstructured 90%% (config 5), none = dense")
    ap.add_argument("--act", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit activation storage; fp16 is what passes the 99.9 %% label gate (DESIGN.md); the line "
                         "always carries the bf16 figure too")
    ap.add_argument("--host-mode", default="pinned", choices=["pinned", "wc", "huge"],
                    help="allocation of the host frame buffers of the e2e legs (drnb200.frameio.HostBuffer)")
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip bf16 / sustained / configs 3-5 / host ceiling")
    ap.add_argument("--layers-out", default=None, help="write the per-layer timing table (JSON) here")
    return ap.parse_args(argv)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_sustained": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0,
            "src": "fallback (B200_PROFILING.md)"}


def build_model(args, device=None):
    """synthetic weights of the named architecture + the masks of the named pruner; returns
    (model, effective state_dict for the oracle, pruner or None)"""
    import numpy as np
    import drnb200
    from drnb200 import synthetic
    model = drnb200.DRNSeg(args.arch, 19, pretrained_model=None, pretrained=False, act_dtype=args.act)
    shapes = collections.OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    sd = synthetic.make_state_dict(shapes, seed=0)
    model.load_state_dict(sd, strict=False)
    pruner = None
    kind = args.pruner if args.sparsity > 0 else "none"
    if kind in ("block", "rmb", "srmbrep"):
        cfg = {"block": lambda: synthetic.block_pruner_config(shapes, args.sparsity),
               "rmb": lambda: synthetic.rmb_pruner_config(shapes, 0.5),
               "srmbrep": lambda: synthetic.srmbrep_config(shapes, args.sparsity)}[kind]()
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
            json.dump(cfg, fh)
        pruner = drnb200.pruners.make_pruner(fh.name, on_gpu=False)
        np.random.seed(0)                                # srmbrep patterns draw from numpy's global RNG
        with contextlib.redirect_stdout(io.StringIO()):  # RmbPruner prints progress like the reference
            if kind == "rmb":
                pruner.generate_masks(model)             # RmbPruner.generate_masks has no is_static (RmbPruner.py:111)
            else:
                pruner.generate_masks(model, is_static=False)   # the reference's magnitude pruning
        os.unlink(fh.name)
        sd = synthetic.sparse_reinit(sd, pruner.mask_dict, seed=0)
        model.load_state_dict(sd, strict=False)
        pruner.apply_masks(model)
    elif kind == "unstructured":
        import torch.nn.utils.prune as prune             # semseg_unstructured.py:769-774: every Conv2d, amount 0.9
        for _, module in model.named_modules():
            if isinstance(module, torch.nn.Conv2d):
                prune.l1_unstructured(module, name="weight", amount=0.9)
        eff = drnb200.checkpoint.normalize_state_dict(model.state_dict())
        eff = eff[0] if isinstance(eff, tuple) else eff
        for k in list(sd):
            if k in eff and k.endswith(".weight") and sd[k].dim() == 4 and not k.startswith("up."):
                sd[k] = eff[k].detach().cpu().clone()
    if device is not None:
        model = model.to(device).eval()
        model.set_pruner(pruner)
    return model, sd, pruner


class ClockSampler(threading.Thread):
    """samples SM clock / power / throttle reasons of one GPU through NVML while a timed region runs"""

    NAMES = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
             "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.power, self.reasons, self.stop_flag, self.max_mhz = [], [], set(), False, None
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
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, b in self.NAMES.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self.stop_flag = True
        self.join()
        return self.summary()

    def summary(self):
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(s)}
        if self.power:
            out["power_w_max"] = round(max(self.power), 1)
        return out


def run_reference(args):
    """the reference's CPU implementation of the path (oracle port: torch CPU fp32, all host threads), timed on
    a bounded sample: one frame per step.  Rank 0 only."""
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
    cfg = workload(args, 1)
    cfg["batch_note"] = ("the CPU arm steps over ONE frame (a bounded sample), the CUDA arm over %d frames per GPU; the "
                         "metric is per frame, so the two values compare directly although the batch sizes differ"
                         % args.batch)
    line = {"impl": "reference", "metric": metric_name(args), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pin_to_gpu_numa_node(local_rank, world):
    """multi-rank runs: bind this process to CPUs near its GPU BEFORE the pinned host buffers are allocated (first
    touch puts their pages on that node).  With the NUMA node of the GPU exposed in sysfs, bind to that node's CPUs;
    otherwise (VMs / containers that show one node) split the affinity mask into `world` contiguous slices by GPU
    index, which at least keeps the ranks' staging threads off each other's cores.  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                       # nvml pads the PCI domain to 8 hex digits
            bus = bus[4:]
        node = -1
        try:
            with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
                node = int(fh.read().strip())
        except OSError:
            pass
        allowed = sorted(os.sched_getaffinity(0))
        if node >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
                cpus = set()
                for part in fh.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= set(allowed)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return "rank bound to NUMA node %d (%d CPUs)" % (node, len(cpus))
        per = max(1, len(allowed) // max(1, world))
        mine = allowed[local_rank * per:(local_rank + 1) * per] or allowed
        os.sched_setaffinity(0, set(mine))
        return "numa_node not exposed; rank bound to CPU slice %d-%d of %d allowed CPUs" % (mine[0], mine[-1], len(allowed))
    except Exception as exc:                                  # topology files missing, no NVML, no permission
        return "not bound (%s)" % type(exc).__name__


def workload(args, batch):
    label = PRUNER_LABEL[args.pruner if args.sparsity > 0 else "none"]
    label = label % (100 * args.sparsity) if "%.0f" in label else label.replace("%%", "%")
    return {"workload": "%s %s, batch %d/GPU, %dx%d, act %s" % (args.arch, label, batch, args.height, args.width, args.act),
            "arch": args.arch, "pruner": args.pruner, "batch_per_gpu": batch, "height": args.height, "width": args.width,
            "sparsity": args.sparsity, "act_dtype": args.act,
            "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2; no explicit flush" % (
                batch * 3 * args.height * args.width * 4 / 1e6)}


class Ctx:
    """process-wide state of one bench run"""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        self.numa = pin_to_gpu_numa_node(self.local, self.world) if self.world > 1 else "single rank: not bound"
        if self.world > 1:
            import torch.distributed as dist
            # NCCL printf()s its version banner to stdout while the communicator is created: keep stdout for the JSON
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = torch.tensor(values, device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather(self, value):
        t = torch.tensor([value], device=self.dev, dtype=torch.float64)
        if self.dist is None:
            return [float(value)]
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(v) for v in out]


def timed_steps(ctx, model, meter, x, gt, steps, warmup):
    """W warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events, max over ranks.
    The evaluation's single collective (all-reduce of the confusion matrix) is inside the timed region and the
    reduced matrix must hold every pixel of every rank (a silently dropped rank fails here)."""
    def step():
        labels = model.predict(x)
        meter.update(labels, gt)
        return labels

    for _ in range(max(3, warmup)):
        step()
    if ctx.dist is not None:
        meter.all_reduce()          # warm NCCL
    ctx.barrier()
    meter.hist.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        step()
    meter.all_reduce()              # NCCL over NVLink (no-op at N=1)
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    total = int(meter.hist.sum().item())
    want = ctx.world * steps * gt.numel()
    assert total == want, "reduced confusion matrix holds %d pixels, expected %d (world %d)" % (total, want, ctx.world)
    return ctx.max_over_ranks([ms])[0]


def e2e_measure(ctx, model, frames, steps, eval_mode=None, host_mode="pinned"):
    """End to end through the repo's public API, `drnb200.FramePipeline` (the data flow of the reference's video caller):
    every step copies ITS input batch from pinned host memory, runs `DRNSeg.predict` and copies the result back to
    pinned host memory; three batches in flight, copies on their own streams.  `frames`: one host batch (float32 NCHW,
    or uint8 NHWC with set_ingest) that stands for the decoded frames.  The result is the uint8 label map (video mode,
    seg_video.py) or — `eval_mode` = (meter, gt) — the 19x19 confusion matrix accumulated on the device (evaluation
    mode, semantic_seg.py test(): 2.9 kB per step).
    -> (ms for `steps` steps, mean H2D GB/s of this rank's copies, mean D2H GB/s)"""
    import drnb200
    pipe = drnb200.FramePipeline(model, tuple(frames.shape), frames.dtype, output="hist" if eval_mode else "labels",
                                 meter=eval_mode, depth=3, host_mode=host_mode, device=ctx.dev)
    for b in range(pipe.depth):
        pipe.staging(b).copy_(frames)              # "decoded" frames sit in the pinned staging buffers

    def feed(n):
        return (pipe.staging(k) for k in range(n))

    for _ in pipe.run(feed(pipe.depth)):
        pass
    ctx.barrier()
    pipe.record_copies = True
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    got = 0
    for res in pipe.run(feed(steps)):              # every result has reached pinned host memory when it is yielded
        got += 1
    f1.record()
    ctx.barrier()
    wall = time.perf_counter() - t0
    assert got == steps
    ms = max(f0.elapsed_time(f1), 1e3 * wall)
    h2d_gbs, d2h_gbs = pipe.copy_rates()
    pipe.close()
    return ms, h2d_gbs, d2h_gbs


def host_ceiling(ctx, hx, label_bytes, seconds=0.4):
    """Copy rates with ALL ranks copying at once and NO kernel running: every rank moves its frame batch host->device
    and a label-map-sized buffer device->host concurrently, as the e2e pipeline does.  What the host (PCIe roots, IOMMU,
    memory system of the VM) can feed to N GPUs.  Ranks have equal work and the step time is the max over ranks, so
    e2e frames/s cannot exceed N x (slowest rank's H2D rate) / bytes per frame.  -> (h2d GB/s, d2h GB/s) of this rank"""
    dev = ctx.dev
    dst = torch.empty(hx.shape, dtype=hx.dtype, device=dev)
    lab = torch.empty(label_bytes, dtype=torch.uint8, device=dev)
    hlab = torch.empty(label_bytes, dtype=torch.uint8).pin_memory()
    back = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    nbytes = hx.numel() * hx.element_size()

    def one():
        dst.copy_(hx, non_blocking=True)
        with torch.cuda.stream(back):
            hlab.copy_(lab, non_blocking=True)

    for _ in range(2):
        one()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    one()
    e1.record()
    torch.cuda.synchronize()
    reps = max(3, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    reps = int(ctx.max_over_ranks([reps])[0])
    ctx.barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b0.record(back)
    for _ in range(reps):
        one()
    e1.record()
    b1.record(back)
    main.wait_stream(back)
    ctx.barrier()
    return (reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, reps * label_bytes / (b0.elapsed_time(b1) * 1e-3) / 1e9)


def graph_replay_cases(model, x, iters=30):
    """frames/s (wall clock, device synchronised on both sides) of DRNSeg.predict on ONE fixed device buffer: the eager
    launch list vs one CUDA-graph replay (DRNSeg.enable_graphs).  Cases: the benchmark batch, one full-size frame (the
    reference's test() loop feeds one frame per step) and one 304x304 frame (seg_video_old.py:127 resizes to 300x300)."""
    N, _, H, W = x.shape
    cases = [("batch%d_%dx%d" % (N, H, W), x), ("batch1_%dx%d" % (H, W), x[:1].contiguous()),
             ("batch1_304x304", x[:1, :, :304, :304].contiguous())]
    out = {}
    for name, xin in cases:
        row = {}
        for mode in ("eager", "graphs"):
            model.enable_graphs(mode == "graphs")
            for _ in range(3):
                model.predict(xin)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                model.predict(xin, static_output=True)
            torch.cuda.synchronize()
            row[mode] = round(xin.shape[0] * iters / (time.perf_counter() - t0), 1)
        out[name] = row
    model.enable_graphs(False)
    return out


def layer_table(eng, per_layer, B, H, W, pk):
    """per conv launch: live MACs, algorithmic bytes, achieved TFLOP/s and GB/s, and the fraction of whichever
    roofline (HBM at the measured copy bandwidth, tensor at the measured burst peak) bounds that launch"""
    import drnb200
    lib = drnb200.ffi.lib()
    KERNELS = {0: "conv_tc<T>", 1: "conv_tc<P>", 2: "conv_tc<T,f32>", 3: "conv_gather", 4: "conv_halo",
               5: "conv_tc<T,ROW>", 6: "conv_tc<T,ROW,PIX>", 7: "conv_ty", 8: "conv_s2", 9: "conv_ys", 10: "conv_y2", -1: "conv_direct"}
    layers = []
    shapes = {-1: (H, W)}
    ops = eng.last_ops or eng.ops
    for i, op in enumerate(ops):
        src = op.input_from if op.input_from is not None else i - 1
        oh, ow = op.out_hw(*shapes[src])
        shapes[i] = (oh, ow)
        macs = B * oh * ow * op.live_elems
        c = op.conv
        io_bytes = 2 * B * (shapes[src][0] * shapes[src][1] * c.in_channels + oh * ow * op.out_channels)
        if op.residual_from is not None:          # residual (or projection input) read
            rsrc = ops[op.residual_from].out_channels if op.residual_from >= 0 else eng.stem[0].out_channels
            io_bytes += 2 * B * oh * ow * (getattr(op, "proj_cin", 0) or min(rsrc, c.out_channels))
        lms = per_layer[op.key]
        mode = max([lib.drnb200_conv_plan_mode(pl) for pl in op.plans.values()] or [-1])
        t_hbm = io_bytes / (pk["hbm_gbs"] * 1e9) * 1e3
        t_tc = 2 * macs / (pk["tflops_burst"] * 1e12) * 1e3
        layers.append({"layer": op.key, "ms": lms, "live_gmac": macs / 1e9, "kernel": KERNELS.get(mode, "?"),
                       "tflops_live": 2 * macs / (lms * 1e-3) / 1e12,
                       "tensor_frac_burst": 2 * macs / (lms * 1e-3) / 1e12 / pk["tflops_burst"],
                       "bytes": io_bytes, "hbm_gbs": io_bytes / (lms * 1e-3) / 1e9,
                       "hbm_frac": io_bytes / (lms * 1e-3) / 1e9 / pk["hbm_gbs"],
                       "bound": "hbm" if t_hbm > t_tc else "tensor", "floor_ms": max(t_hbm, t_tc),
                       "frac_of_bound": max(t_hbm, t_tc) / lms,
                       "live_tiles": op.n_live, "tile": [op.tile_o, op.tile_ci]})
    return layers


def per_layer_pass(eng, x, reps=3):
    per_layer = collections.OrderedDict()
    for _ in range(reps):
        tl = []
        eng.run(x, want_labels=True, timings=tl)
        torch.cuda.synchronize()
        for name, a, b in tl:
            per_layer[name] = per_layer.get(name, 0.0) + a.elapsed_time(b) / reps
    return per_layer


def short_config(ctx, base_args, arch, pruner, batch, steps=5):
    """one of BASELINE configs 3/4/5 on its own masks: device-timed frames/s + the all-conv tensor fraction"""
    import drnb200
    args = argparse.Namespace(**vars(base_args))
    args.arch, args.pruner, args.batch = arch, pruner, batch
    t0 = time.perf_counter()
    model, _, _ = build_model(args, ctx.dev)
    build_s = time.perf_counter() - t0
    g = torch.Generator(device=ctx.dev).manual_seed(4321 + ctx.rank)
    x = torch.randn(batch, 3, args.height, args.width, device=ctx.dev, generator=g)
    gt = torch.randint(0, 19, (batch, args.height, args.width), device=ctx.dev, generator=g,
                       dtype=torch.int64).to(torch.uint8)
    meter = drnb200.ConfusionMeter(19, ctx.dev)
    with torch.no_grad():
        ms = timed_steps(ctx, model, meter, x, gt, steps, 3)
        eng = model.engine()
        per_layer = per_layer_pass(eng, x, reps=2)
    pk = peaks()
    dense, live, tile = eng.mac_counts(batch, args.height, args.width)
    conv_ms = sum(v for k, v in per_layer.items() if k not in ("stem", "head"))
    conv_live = live - batch * args.height * args.width * 16 * 147 \
        - batch * (args.height // 8) * (args.width // 8) * model.seg.in_channels * 19
    conv_tf = 2.0 * conv_live / (conv_ms * 1e-3) / 1e12
    out = {"config": workload(args, batch), "value": ctx.world * batch * steps / (ms * 1e-3), "unit": "frames/s",
           "ms_per_step": ms / steps, "steps": steps, "launches_per_step": eng.launches_per_forward + 1,
           "all_convs_tflops_unpruned": conv_tf, "all_convs_frac_of_burst": conv_tf / pk["tflops_burst"],
           "macs_per_frame": {"dense_g": dense / batch / 1e9, "unpruned_g": live / batch / 1e9,
                              "live_tile_g": tile / batch / 1e9},
           "model_build_s": round(build_s, 1)}
    model._close_engines()
    del model, x, gt
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import drnb200
    from drnb200 import synthetic
    from drnb200.frameio import HostBuffer
    ctx = Ctx()
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    model, sd, pruner = build_model(args, dev)
    B, H, W = args.batch, args.height, args.width
    # synthetic frames: a distinct batch per rank, generated on the device outside the timed region
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, 3, H, W, device=dev, generator=g)
    gt = torch.randint(0, 19, (B, H, W), device=dev, generator=g, dtype=torch.int64).to(torch.uint8)
    meter = drnb200.ConfusionMeter(19, dev)
    eng = model.engine()
    extras = not args.no_extras
    pk = peaks()

    with torch.no_grad():
        sampler = ClockSampler(ctx.local)
        sampler.start()
        ms = timed_steps(ctx, model, meter, x, gt, args.steps, args.warmup)
        launches = (eng.launches_per_forward + 1) * args.steps
        miou = meter.miou()

        # ---- e2e: host frames -> H2D -> predict -> D2H labels, through the public API (drnb200.FramePipeline)
        hbuf = HostBuffer((B, 3, H, W), torch.float32, args.host_mode)
        hbuf.tensor.copy_(x.cpu())
        e2e_ms, h2d_gbs, d2h_gbs = e2e_measure(ctx, model, hbuf.tensor, args.steps, host_mode=args.host_mode)
        ceiling, ceiling_d2h = host_ceiling(ctx, hbuf.tensor, B * H * W) if extras else (None, None)
        # the same through the fused frame ingest: uint8 HWC frames as cv2 delivers them, the
        # reference's ToTensor + Normalize (info.json statistics) applied inside the stem kernel
        e2e_u8_ms = u8_h2d = u8_d2h = e2e_ev_ms = ev_h2d = None
        if W % 16 == 0:
            model.set_ingest(INFO_MEAN, INFO_STD)
            hu8 = HostBuffer((B, H, W, 3), torch.uint8, args.host_mode)
            hu8.tensor.copy_(synthetic.make_u8_frames(B, H, W, seed=1234 + rank))
            e2e_u8_ms, u8_h2d, u8_d2h = e2e_measure(ctx, model, hu8.tensor, args.steps, host_mode=args.host_mode)
            if extras:
                e2e_ev_ms, ev_h2d, _ = e2e_measure(ctx, model, hu8.tensor, args.steps, eval_mode=(meter, gt),
                                                   host_mode=args.host_mode)
            hu8.close()
        clocks = sampler.finish()
        e2e_ms, e2e_u8_ms, e2e_ev_ms = ctx.max_over_ranks([e2e_ms, e2e_u8_ms or 0.0, e2e_ev_ms or 0.0])
        copy_rates = {k: ctx.gather(v) for k, v in (("h2d", h2d_gbs), ("d2h", d2h_gbs), ("u8_h2d", u8_h2d or 0.0),
                                                    ("u8_d2h", u8_d2h or 0.0), ("ev_h2d", ev_h2d or 0.0),
                                                    ("ceiling", ceiling or 0.0), ("ceiling_d2h", ceiling_d2h or 0.0))}
        hbuf.close()

        # ---- per-layer timing pass (CUDA events per launch) for the roofline objects
        per_layer = per_layer_pass(eng, x)
        layers = layer_table(eng, per_layer, B, H, W, pk)

        # ---- sustained: >= N seconds of back-to-back steps (clocks, power and throttle reasons under real load)
        sustained = None
        if extras and args.sustained_seconds > 0:
            n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / (ms / args.steps)) + 1)
            s2 = ClockSampler(ctx.local, period=0.1)
            s2.start()
            sus_ms = timed_steps(ctx, model, meter, x, gt, n_sus, 3)
            sustained = dict(s2.finish(), seconds=sus_ms * 1e-3, steps=n_sus, value=world * B * n_sus / (sus_ms * 1e-3),
                             unit="frames/s", ms_per_step=sus_ms / n_sus)

        # ---- bf16 activation storage on the same workload (north_star's nominal layout), beside the fp16 headline
        other = None
        labels_main = model.predict(x[:1]).clone()
        if extras:
            odt = "bf16" if args.act == "fp16" else "fp16"
            model.set_act_dtype(odt)
            o_ms = timed_steps(ctx, model, meter, x, gt, args.steps, 3)
            agree = float((model.predict(x[:1]) == labels_main).float().mean())
            other = {"act_dtype": odt, "value": world * B * args.steps / (o_ms * 1e-3), "unit": "frames/s",
                     "ms_per_step": o_ms / args.steps, "label_agreement_with_%s_path" % args.act: agree}
            model.set_act_dtype(args.act)
        graphs = graph_replay_cases(model, x) if extras and H >= 304 and W >= 304 else None

    if rank == 0:
        frames = world * B * args.steps
        dense, live, tile = eng.mac_counts(B, H, W)
        conv_ms = sum(v for k, v in per_layer.items() if k not in ("stem", "head"))
        conv_live = live - B * H * W * 16 * 147 - B * (H // 8) * (W // 8) * model.seg.in_channels * 19
        conv_tflops = 2.0 * conv_live / (conv_ms * 1e-3) / 1e12
        head_bytes = B * ((H // 8) * (W // 8) * model.seg.in_channels * 2 + H * W) + 19 * model.seg.in_channels * 4 + 19 * 4
        head_gbs = head_bytes / (per_layer["head"] * 1e-3) / 1e9
        # the dominant kernel = the one with the largest share of the step (conv_tc_kernel<MODE_T, ROW>: the 3x3
        # stride-1 convs of layers 4-8), aggregated over its launches of one step
        by_kernel = collections.defaultdict(float)
        fam = lambda name: name.replace(",PIX", "")     # noqa: E731  (both accumulator layouts of the row-halo kernel)
        for l in layers:
            by_kernel[fam(l["kernel"])] += l["ms"]
        dom_name = max(by_kernel, key=by_kernel.get)
        dom = [l for l in layers if fam(l["kernel"]) == dom_name]
        dom_ms = sum(l["ms"] for l in dom)
        dom_tflops = 2.0 * sum(l["live_gmac"] for l in dom) * 1e9 / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0
        dom_floor = sum(l["floor_ms"] for l in dom)
        # DRAM traffic per launch from the committed `ncu --set full` capture of the same workload
        # (tools/ncu_summarize.py traffic -> profiles/ncu_traffic.json); null when there is none for this workload
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = {k: v for k, v in json.load(fh).items() if v.get("workload") == workload(args, B)["workload"]}
        dom_traffic = traffic.get("conv_tc_row", {}).get("bytes_per_launch") if "ROW" in dom_name else None
        dom_io = sum(l["bytes"] for l in dom) / max(len(dom), 1)    # algorithmic bytes per launch
        if args.layers_out:
            with open(args.layers_out, "w") as fh:
                json.dump({"per_layer_ms": per_layer, "layers": layers, "batch": B}, fh, indent=1)
        frame_bytes_f32, frame_bytes_u8, label_bytes = 3 * H * W * 4, 3 * H * W, H * W
        ceil_rates = [v for v in copy_rates["ceiling"] if v > 0]
        line = {
            "metric": metric_name(args), "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.act,
            "data": "synthetic", "config": workload(args, B),
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": B * frame_bytes_f32, "d2h_bytes_per_step": B * label_bytes,
                    "h2d_gbs_per_rank": [round(v, 2) for v in copy_rates["h2d"]],
                    "d2h_gbs_per_rank": [round(v, 2) for v in copy_rates["d2h"]],
                    "host_buffers": args.host_mode},
            "e2e_uint8_frames": None if not e2e_u8_ms else {
                "value": frames / (e2e_u8_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes_u8,
                "d2h_bytes_per_step": B * label_bytes,
                "h2d_gbs_per_rank": [round(v, 2) for v in copy_rates["u8_h2d"]],
                "d2h_gbs_per_rank": [round(v, 2) for v in copy_rates["u8_d2h"]],
                "note": "same pipeline fed with uint8 HWC frames; ToTensor+Normalize fused into the stem kernel"},
            "gpu_launches": launches,
            "graph_replay": None if graphs is None else dict(
                graphs, unit="frames/s", note="DRNSeg.predict on one fixed device buffer, wall clock: eager launch list "
                "vs one CUDA-graph replay (enable_graphs); the headline numbers above use the eager list"),
            "host_numa": ctx.numa,
            "clocks": clocks,
            "roofline": {"bound": "tensor",
                         "kernel": "%s (%d launches per step: %s; * = pixel-major accumulator flavour)" % (
                             dom_name, len(dom), ", ".join(l["layer"].replace("layer.", "") + ("*" if "PIX" in l["kernel"] else "")
                                                           for l in dom)),
                         "achieved": dom_tflops, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                         "frac": dom_tflops / pk["tflops_burst"],
                         "frac_of_sustained_peak": dom_tflops / pk["tflops_sustained"],
                         "frac_of_per_launch_bound": dom_floor / dom_ms if dom_ms else None,
                         "traffic": dom_traffic,
                         "traffic_note": "avg DRAM bytes read+written per launch (ncu --set full, profiles/ncu_traffic.json)"
                                         "; algorithmic input+output+residual bytes per launch: %.0f" % dom_io,
                         "peak_source": pk["src"] + ": bf16_tflops (burst) - each launch is event-timed on its own; "
                                        "frac_of_sustained_peak uses bf16_tflops_sustained",
                         "per_launch_bound_note": "launches of layers 4-5 move more bytes than their MACs cover: their floor "
                                                  "is HBM (hbm_gbs), not the tensor pipe; frac_of_per_launch_bound = sum of "
                                                  "per-launch max(HBM floor, tensor floor) / measured",
                         "flops": "2 x unpruned (mask != 0) MACs of these launches",
                         "ms_per_step": dom_ms},
            "roofline_all_convs": {"bound": "tensor", "kernel": "all %d conv launches of one step (stem/seg excluded)" % len(layers),
                                   "achieved": conv_tflops, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                                   "frac": conv_tflops / pk["tflops_burst"],
                                   "frac_of_sustained_peak": conv_tflops / pk["tflops_sustained"],
                                   "frac_of_per_launch_bound": sum(l["floor_ms"] for l in layers) / conv_ms,
                                   "ms_per_step": conv_ms},
            "roofline_head": {"bound": "hbm", "kernel": "head_fused_kernel (classifier GEMM + x8 upsample + argmax, one launch)",
                              "achieved": head_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": head_gbs / pk["hbm_gbs"],
                              "traffic": traffic.get("head_fused", {}).get("bytes_per_launch"),
                              "bytes_per_launch": head_bytes, "ms": per_layer["head"]},
            "stem_ms_per_step": per_layer["stem"],
            "macs_per_frame": {"dense_g": dense / B / 1e9, "unpruned_g": live / B / 1e9,
                               "live_tile_g": tile / B / 1e9},
            "miou_vs_random_labels": miou,
            "reduced_histogram_pixels": "asserted == world x steps x batch x H x W in every timed region",
        }
        if e2e_ev_ms:
            line["e2e_uint8_eval_mode"] = {
                "value": frames / (e2e_ev_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes_u8,
                "d2h_bytes_per_step": 19 * 19 * 8, "h2d_gbs_per_rank": [round(v, 2) for v in copy_rates["ev_h2d"]],
                "note": "evaluation flow of semantic_seg.py test(): uint8 frames in, confusion matrix updated on the device, "
                        "the 19x19 int64 matrix read back every step instead of the label map"}
        if ceil_rates:
            slow = min(ceil_rates)
            dev_rate = frames / (ms * 1e-3)
            b_f32 = min(world * slow * 1e9 / frame_bytes_f32, dev_rate)
            b_u8 = min(world * slow * 1e9 / frame_bytes_u8, dev_rate)
            line["host_ceiling"] = {
                "what": "copy rates with all %d ranks moving a frame batch H2D and a label map D2H at once, no kernels; ranks "
                        "have equal work and a step ends with the slowest rank, so e2e <= N x slowest H2D rate / bytes "
                        "per frame (and <= the device-timed rate)" % world,
                "h2d_gbs_per_rank": [round(v, 2) for v in ceil_rates], "h2d_gbs_sum": round(sum(ceil_rates), 1),
                "h2d_gbs_slowest_rank": round(slow, 2),
                "d2h_gbs_per_rank": [round(v, 2) for v in copy_rates["ceiling_d2h"]],
                "e2e_bound_f32_frames": b_f32, "e2e_bound_uint8_frames": b_u8,
                "e2e_over_bound": (frames / (e2e_ms * 1e-3)) / b_f32,
                "e2e_uint8_over_bound": None if not e2e_u8_ms else (frames / (e2e_u8_ms * 1e-3)) / b_u8,
                "e2e_uint8_eval_mode_over_bound": None if not e2e_ev_ms else (frames / (e2e_ev_ms * 1e-3)) / b_u8}
        if sustained:
            line["sustained"] = sustained
        if other:
            line[other["act_dtype"]] = other

    # ---- BASELINE configs 3 / 4 / 5 on their own masks (short; all ranks take part so that N>1 runs shard them too)
    sub = []
    if extras and args.arch == "drn_d_22" and args.pruner == "block" and (H, W) == (1024, 2048):
        with torch.no_grad():
            for arch, prn, bsz in (("drn_d_38", "rmb", 8), ("drn_d_54", "srmbrep", 4), ("drn_d_22", "unstructured", 8),
                                   ("drn_d_22", "none", 8)):
                a2 = argparse.Namespace(**vars(args))
                if prn == "none":
                    a2.sparsity = 0.0
                    prn = "block"
                sub.append(short_config(ctx, a2, arch, prn, bsz))
    if rank == 0:
        if sub:
            dense_fps = sub[-1]["value"]
            line["configs"] = sub
            line["speedup_over_same_kernels_dense"] = {
                "block_75": line["value"] / dense_fps, "unstructured_90": sub[2]["value"] / dense_fps,
                "note": "frames/s of the masked network / frames/s of the same kernels with no masks (config 5's question)"}
        if not args.no_cpu_baseline and world == 1:      # reported at N=1 only (rank 0 owns the host cores there)
            from oracle import drn_oracle           # cpu_baseline leg: the oracle port, bounded sample; also the checker
            xs = synthetic.make_frames(1, H, W, seed=1234)
            with torch.no_grad():
                ref_lp, ref_seg = drn_oracle.drnseg_forward(sd, xs)
                ref_lab = torch.max(ref_lp, 1)[1]
                t0 = time.perf_counter()
                n_it = 3
                for _ in range(n_it):
                    torch.max(drn_oracle.drnseg_forward(sd, xs)[0], 1)
                dt = (time.perf_counter() - t0) / n_it
                par = {"frame": "1 x %dx%d torch.randn frame (seed 1234) against the fp32 oracle" % (H, W)}
                for adt in ("fp16", "bf16"):
                    model.set_act_dtype(adt)
                    lab = model.predict(xs.to(dev)).cpu().long()
                    seg = model(xs.to(dev))[1].cpu()
                    par[adt] = {"label_agreement": float((lab == ref_lab).float().mean()),
                                "logits_rel_err": float((seg - ref_seg).abs().max() / ref_seg.abs().max())}
                model.set_act_dtype(args.act)
            line["parity"] = par
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": "1 frame %dx%d x %d iterations, torch CPU fp32 (oracle port of "
                                              "semantic_seg.DRNSeg + torch.max)" % (H, W, n_it)}
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
