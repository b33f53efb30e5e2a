#!/usr/bin/env python
"""Storage-precision error budget of the 16-bit activation path, emulated on the CPU with the oracle.

The CUDA path accumulates in fp32 and rounds ONCE per stored tensor (frame, every conv+BN(+res)+ReLU output, every
stored shortcut); this script replays exactly those roundings inside the fp32 oracle (`quant` hook of
oracle/drn_oracle.py) and reports label agreement / logit error against the plain fp32 run, for:
  * fp16 and bf16 storage everywhere (what the engine does),
  * bf16 with ONE group of tensors kept in fp32 (which roundings cost the margin),
  * bf16 conv operands with an fp32 residual stream + fp32 hand-off into the head (SURVEY 7.3-2 iii: the bound on
    what a wider residual stream could buy).
Test infrastructure (imports oracle/); writes profiles/r02_precision_budget.txt.

  python tests/precision_budget.py [--arch drn_d_22] [--hw 256 512] [--seed 5] [--dense]
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import conftest  # noqa: F401,E402  (puts the repo root and the package on sys.path)
import torch  # noqa: E402

from helpers import gate_case_cpu  # noqa: E402
from oracle import drn_oracle, recipe  # noqa: E402


def rounder(dtype):
    return lambda t: t.to(dtype).to(torch.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="drn_d_22")
    ap.add_argument("--hw", type=int, nargs=2, default=(256, 512))
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--out", default=os.path.join(conftest.ROOT, "profiles", "r02_precision_budget.txt"))
    args = ap.parse_args()
    _, sd, _ = gate_case_cpu(args.arch, not args.dense, args.seed)
    x = recipe.make_frames(2, args.hw[0], args.hw[1], seed=99 + args.seed)
    ref_lp, ref_seg = drn_oracle.drnseg_forward(sd, x)
    ref_lab = ref_lp.argmax(1)
    top2 = ref_lp.topk(2, dim=1)[0]
    margin = (top2[:, 0] - top2[:, 1])
    rng = float(ref_seg.abs().max())

    def run(name, quant):
        lp, seg = drn_oracle.drnseg_forward(sd, x, quant=quant)
        agree = float((lp.argmax(1) == ref_lab).float().mean())
        err = float((seg - ref_seg).abs().max()) / rng
        rms = float((seg - ref_seg).pow(2).mean().sqrt()) / rng
        return "%-58s labels %.5f   logits max %.2e  rms %.2e" % (name, agree, err, rms)

    f16, b16 = rounder(torch.float16), rounder(torch.bfloat16)
    block_out = lambda key: key.endswith(".conv2") and ".conv3" not in key or key.endswith(".conv3")  # noqa: E731
    lines = ["# storage-precision budget: %s, %s, %dx%d, seed %d (CPU emulation of the engine's roundings)" % (
        args.arch, "dense" if args.dense else "BlockPruner 75 %", args.hw[0], args.hw[1], args.seed),
        "# fp32 logit range %.3f; median top-1/top-2 log-prob margin %.4f; pixels with margin < 1e-2*range: %.2f %%" % (
            rng, float(margin.median()), 100 * float((margin < 1e-2 * rng).float().mean()))]
    lines.append(run("fp16 everywhere (engine act_dtype=fp16)", lambda r, k, t: f16(t)))
    lines.append(run("bf16 everywhere (engine act_dtype=bf16)", lambda r, k, t: b16(t)))
    lines.append(run("bf16, frame kept fp32", lambda r, k, t: t if r == "input" else b16(t)))
    lines.append(run("bf16, stored shortcuts kept fp32", lambda r, k, t: t if r == "shortcut" else b16(t)))
    lines.append(run("bf16, layer-8 output kept fp32 (fp32 hand-off to the head)",
                     lambda r, k, t: t if k.startswith("layer.8.") else b16(t)))
    lines.append(run("bf16, layers 7+8 kept fp32",
                     lambda r, k, t: t if k.startswith(("layer.7.", "layer.8.")) else b16(t)))
    lines.append(run("bf16, front (frame, stem, layers 1-3) kept fp32",
                     lambda r, k, t: t if r == "input" or k.startswith(("layer.0.", "layer.1.", "layer.2.", "layer.3."))
                     else b16(t)))
    lines.append(run("bf16 conv1 outputs, block outputs (residual stream) fp16",
                     lambda r, k, t: f16(t) if (r == "act" and block_out(k)) or r == "shortcut" else b16(t)))

    # fp32 residual stream: block outputs stay fp32 in "HBM", every conv reads a bf16 rounding of its input
    class Stream:
        """quant hook that keeps the stored tensor in fp32 but hands the NEXT conv a bf16 copy: emulated by rounding
        conv inputs instead of outputs, i.e. rounding in a forward pre-hook of F.conv2d"""
    import torch.nn.functional as F
    real_conv = F.conv2d

    def conv_bf16_inputs(inp, w, b=None, **kw):
        return real_conv(b16(inp), w, b, **kw)
    F.conv2d = conv_bf16_inputs
    try:
        lines.append(run("bf16 conv operands, fp32 residual stream + fp32 head hand-off", None))
    finally:
        F.conv2d = real_conv
    text = "\n".join(lines)
    print(text)
    with open(args.out, "a") as fh:
        fh.write(text + "\n\n")


if __name__ == "__main__":
    main()
