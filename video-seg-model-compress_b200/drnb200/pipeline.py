"""Streaming frame pipeline: the data flow of the reference's video caller as one object.

``FrameCapture`` (seg_video_old.py:110-203, seg_video_new.py:112-176) takes decoded frames, resizes each with
``T.Resize`` on the CPU, normalises them into an fp32 batch, copies it to the GPU, runs ``model(img)[0]``,
``torch.max(final, 1)``, copies the label maps back and draws ``CITYSCAPE_PALETTE[pred]`` over the frame with
alpha 0.6.  Here every step between "decoded frames in host memory" and "label maps / overlays in host memory" runs on
the device and the three stages overlap:

    host batch k+1  --H2D (copy stream)-->  device frames
    device frames k --[resize_frames] -> DRNSeg.predict (fused uint8 ingest) -> [overlay]-->  device result
    device result k-1 --D2H (copy-back stream)-->  pinned host result

``FramePipeline.run(batches)`` yields one host result per input batch, in order; every batch is copied in from pinned
host memory and its result copied back, nothing is cached.  With ``meter=(ConfusionMeter, labels)`` the confusion
matrix is accumulated on the device instead (evaluation flow of semantic_seg.py test(): 2.9 kB per batch come back).
This is the call ``bench.py`` times as ``e2e``.  There is no CPU path.
"""
import torch

from . import ffi
from . import frameio


class FramePipeline:
    """model: drnb200.DRNSeg on a CUDA device (``set_ingest`` done when frames are uint8).

    frame_shape / dtype: shape of ONE host batch — uint8 ``[B,H,W,3]`` (HWC, as cv2 / PIL deliver frames) or float32
    ``[B,3,H,W]`` (already normalised, the reference's tensor).  resize_to=(h, w): PIL-exact resize on the device first
    (uint8 frames only; w % 16 == 0).  output: "labels" (uint8 [B,h,w]), "overlay" (uint8 [B,h,w,3], palette blended
    over the frame with `alpha`, seg_video.py:200-203) or "hist" (int64 [classes, classes], needs `meter`).
    depth: batches in flight (input / output buffers); host_mode: drnb200.frameio.HostBuffer mode of the input staging
    buffers ("pinned", "wc" or "huge")."""

    def __init__(self, model, frame_shape, dtype=torch.uint8, resize_to=None, output="labels", alpha=0.6, meter=None,
                 depth=3, host_mode="pinned", device=None):
        if output not in ("labels", "overlay", "hist"):
            raise ffi.Drnb200Error("FramePipeline output must be 'labels', 'overlay' or 'hist'")
        if output == "hist" and meter is None:
            raise ffi.Drnb200Error("output='hist' needs meter=(ConfusionMeter, ground-truth labels on the device)")
        if dtype not in (torch.uint8, torch.float32):
            raise ffi.Drnb200Error("frames must be uint8 [B,H,W,3] or float32 [B,3,H,W]")
        if dtype != torch.uint8 and (resize_to is not None or output == "overlay"):
            raise ffi.Drnb200Error("resize_to / overlay need uint8 frames")
        self.model, self.output, self.alpha, self.meter = model, output, float(alpha), meter
        self.resize_to = None if resize_to is None else (int(resize_to[0]), int(resize_to[1]))
        self.dev = torch.device(device) if device is not None else next(model.parameters()).device
        if self.dev.type != "cuda":
            raise ffi.Drnb200Error("FramePipeline needs a model on a CUDA device; there is no CPU path")
        self.depth = max(2, int(depth))
        self.shape, self.dtype = tuple(int(v) for v in frame_shape), dtype
        B = self.shape[0]
        h, w = (self.shape[1], self.shape[2]) if dtype == torch.uint8 else (self.shape[2], self.shape[3])
        if self.resize_to is not None:
            h, w = self.resize_to
        oh, ow = 8 * (-(-h // 8)), 8 * (-(-w // 8))            # the label map follows the reference's rounding
        with torch.cuda.device(self.dev):
            self.copy_in = torch.cuda.Stream(device=self.dev)
            self.copy_out = torch.cuda.Stream(device=self.dev)
            self.h_in = [frameio.HostBuffer(self.shape, dtype, host_mode) for _ in range(self.depth)]
            self.d_in = [torch.empty(self.shape, dtype=dtype, device=self.dev) for _ in range(self.depth)]
            # resized frames live in per-slot buffers too: fixed addresses are what CUDA-graph replay keys on
            self.d_rs = [torch.empty((B, h, w, 3), dtype=torch.uint8, device=self.dev) for _ in range(self.depth)] \
                if self.resize_to is not None else None
            if output == "labels":
                oshape, odt = (B, oh, ow), torch.uint8
            elif output == "overlay":
                if (oh, ow) != (h, w):
                    raise ffi.Drnb200Error("overlay needs frame sizes that are multiples of 8 (got %dx%d)" % (h, w))
                oshape, odt = (B, oh, ow, 3), torch.uint8
            else:
                oshape, odt = tuple(meter[0].hist.shape), torch.int64
            self.h_out = [torch.empty(oshape, dtype=odt).pin_memory() for _ in range(self.depth)]
            self.ready = [torch.cuda.Event() for _ in range(self.depth)]       # input k is on the device
            self.consumed = [torch.cuda.Event() for _ in range(self.depth)]    # kernels of batch k have read their input
            self.returned = [torch.cuda.Event() for _ in range(self.depth)]    # result k is in host memory
            self.keep = [None] * self.depth
            main = torch.cuda.current_stream(self.dev)
            for b in range(self.depth):
                self.consumed[b].record(main)
                self.returned[b].record(main)
        self.record_copies, self._copies = False, []
        self.h2d_bytes = self.h_in[0].nbytes
        self.d2h_bytes = self.h_out[0].numel() * self.h_out[0].element_size()

    # ---- the three stages ------------------------------------------------------------------------------------
    def _stage_in(self, b, batch):
        src = batch if isinstance(batch, torch.Tensor) else torch.from_numpy(batch)
        if tuple(src.shape) != self.shape or src.dtype != self.dtype:
            raise ffi.Drnb200Error("FramePipeline was built for %s %s batches, got %s %s" % (
                self.dtype, self.shape, src.dtype, tuple(src.shape)))
        buf = self.h_in[b].tensor
        if src.data_ptr() != buf.data_ptr():           # a caller may also decode straight into staging(b) and pass it
            self.ready[b].synchronize()                # the previous H2D copy out of this staging buffer has finished
            buf.copy_(src)
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.consumed[b])  # the kernels that read the device copy of slot b are done
            t0 = self._mark(self.copy_in)
            self.d_in[b].copy_(buf, non_blocking=True)
            self._mark(self.copy_in, t0, "h2d")
            self.ready[b].record(self.copy_in)

    def _mark(self, stream, start=None, kind=None):
        """optional CUDA-event timing of the copies (record_copies=True): -> copy_rates()"""
        if not self.record_copies:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        if start is not None:
            self._copies.append((kind, start, ev))
        return ev

    def copy_rates(self):
        """(mean H2D GB/s, mean D2H GB/s) of the copies timed since record_copies was switched on"""
        torch.cuda.synchronize(self.dev)
        out = []
        for kind, nbytes in (("h2d", self.h2d_bytes), ("d2h", self.d2h_bytes)):
            ms = [a.elapsed_time(b) for k, a, b in self._copies if k == kind]
            out.append(nbytes / (sum(ms) / len(ms) * 1e-3) / 1e9 if ms and sum(ms) > 0 else 0.0)
        self._copies = []
        return tuple(out)

    def staging(self, b):
        """pinned input buffer of slot b % depth, safe to overwrite (its previous H2D copy has finished): decode the
        next batch straight into it and pass it to run() to skip the host-side copy"""
        self.ready[b % self.depth].synchronize()
        return self.h_in[b % self.depth].tensor

    @torch.no_grad()
    def _compute(self, b):
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self.ready[b])
        main.wait_event(self.returned[b])              # the device result of the batch that used this slot is on the host
        x = self.d_in[b]
        if self.resize_to is not None:
            x = frameio.resize_frames(x, self.resize_to, out=self.d_rs[b])
        # with model.enable_graphs() the forward is one graph launch per staging buffer (d_in[b] is a fixed address); the
        # graph's own label tensor is consumed by the copy-back / overlay / histogram right behind it on this stream
        labels = self.model.predict(x, static_output=True) if getattr(self.model, "use_graphs", False) \
            else self.model.predict(x)
        if self.output == "overlay":
            res = frameio.overlay(labels, x, self.alpha)
        elif self.output == "hist":
            self.meter[0].update(labels, self.meter[1])
            res = self.meter[0].hist
        else:
            res = labels
        self.keep[b] = res
        self.consumed[b].record(main)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.consumed[b])
            t0 = self._mark(self.copy_out)
            self.h_out[b].copy_(res, non_blocking=True)
            self._mark(self.copy_out, t0, "d2h")
            self.returned[b].record(self.copy_out)

    def run(self, batches):
        """batches: iterable of host batches (numpy arrays or CPU tensors of the configured shape).  Yields the pinned
        host result of every batch, in order; a yielded tensor is valid until `depth` more batches have been fed."""
        with torch.cuda.device(self.dev):
            k = done = 0
            for batch in batches:
                # slot k % depth is about to be reused: the result of its previous occupant (batch k - depth) goes out first
                while done <= k - self.depth:
                    yield self._deliver(done % self.depth)
                    done += 1
                self._stage_in(k % self.depth, batch)
                if k >= 1:
                    self._compute((k - 1) % self.depth)
                k += 1
            if k >= 1:
                self._compute((k - 1) % self.depth)
            while done < k:
                yield self._deliver(done % self.depth)
                done += 1

    def _deliver(self, b):
        self.returned[b].synchronize()
        return self.h_out[b]

    def close(self):
        for hb in self.h_in:
            hb.close()
        self.h_in = []
