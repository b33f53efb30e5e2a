#!/usr/bin/env python
"""Same-box A/B of environment routing knobs, per launch: runs `bench.py --no-extras --layers-out` once per variant
(alternating, `--rounds` times) and prints the per-layer CUDA-event times side by side.
usage: python tools/ab_layers.py "DRNB200_ROW_RING=4,3" "DRNB200_NG=4" [--rounds 2]"""
import json, os, subprocess, sys, tempfile

def main():
    argv, rounds = sys.argv[1:], 2
    if "--rounds" in argv:
        i = argv.index("--rounds")
        rounds = int(argv[i + 1])
        del argv[i:i + 2]
    args = argv
    variants = [("default", {})] + [(a, dict(kv.split("=", 1) for kv in a.split())) for a in args]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {name: [] for name, _ in variants}
    if os.environ.get('AB_DRY'):
        print(variants, rounds); return
    for _ in range(rounds):
        for name, env in variants:
            with tempfile.NamedTemporaryFile(suffix=".json") as f:
                e = dict(os.environ); e.update(env)
                out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-extras", "--steps", "20",
                                      "--warmup", "3", "--layers-out", f.name], env=e, capture_output=True, text=True)
                if out.returncode != 0 or not out.stdout.strip():
                    sys.exit("bench.py failed for variant %r (rc %d):\n%s" % (name, out.returncode, out.stderr[-3000:]))
                line = json.loads(out.stdout.strip().splitlines()[-1])
                res[name].append((line["value"], json.load(open(f.name))["per_layer_ms"]))
    names = [n for n, _ in variants]
    print("%-18s" % "layer" + "".join("%22s" % n[:22] for n in names))
    for layer in res[names[0]][0][1]:
        row = [min(r[1][layer] for r in res[n]) for n in names]
        print("%-18s" % layer + "".join("%14.4f (%+5.1f%%)" % (v, 100 * (v / row[0] - 1)) for v in row))
    print("%-18s" % "frames/s (best)" + "".join("%22.1f" % max(r[0] for r in res[n]) for n in names))

if __name__ == "__main__":
    main()
