"""Label the per-launch GEMM times printed by `bench.py --layers` (stderr) with ResNet50 layer names and floors."""
import sys
TF, BW = 1363.5e12, 6557.8e9
B = 512
CHAIN = "--no-chain" not in sys.argv       # stages 2-3: `_increase` + the next block's `_reduce` are one launch (conv_chain*_kernel)
names, floors = ["conv1"], [max(2 * B * 12544 * 147 * 64 / TF, (B * 115 * 115 * 16 + B * 12544 * 64) * 2 / BW) * 1e6]
cin = 64
for si, (nb, mid, H) in enumerate([(3, 64, 56), (4, 128, 28), (6, 256, 14), (3, 512, 7)]):
    cout = mid * 4
    chained_in = False                     # this block's reduce was computed by the previous block's chained launch
    for b in range(nb):
        M = B * H * H
        def add(n, flops, elems):
            names.append("s%db%d %s" % (si + 2, b + 1, n)); floors.append(max(flops / TF, elems * 2 / BW) * 1e6)
        if not chained_in:
            add("reduce", 2 * M * cin * mid, M * cin + M * mid)
        add("3x3", 2 * M * 9 * mid * mid, 2 * M * mid)
        if b == 0: add("proj", 2 * M * cin * cout, M * cin + M * cout)
        chain_out = CHAIN and mid <= 128 and b + 1 < nb
        if chain_out:   # reads T2 + residual, writes the block output and the next block's T1; the next reduce's input never leaves the chip / L2
            add("incr+red", 2 * M * mid * cout + 2 * M * cout * mid, M * mid + 2 * M * cout + M * mid)
        else:
            add("incr", 2 * M * mid * cout, M * mid + 2 * M * cout)
        chained_in = chain_out
        cin = cout
for path in [a for a in sys.argv[1:] if not a.startswith("--")]:
    lines = [l for l in open(path) if l.strip() and l.strip()[0].isdigit()]
    vals = [float(x) for x in lines[-1].split()]
    n = len(names)
    head = 12 if (len(vals) - 12) % n == 0 else 0                # e2e steps end with the 12 PhaseNet launches; --config resnet512 has none
    passes = max(1, (len(vals) - head) // n)                     # ResNet passes per step (images per step / images per pass)
    step_images = 2048 if head else 512
    B = step_images // passes
    chunk = [sum(vals[c * n + i] for c in range(passes)) / passes * (512.0 / B) for i in range(n)]      # per 512 images
    print("== %s: %d launches per step, %d-image ResNet passes; per 512 images: %.0f us measured, %.0f us floor (sum of per-layer max(tensor, HBM))" % (path, len(vals), B, sum(chunk), sum(floors)))
    for nm, v, f in zip(names, chunk, floors):
        print("  %-12s %7.1f us  floor %6.1f  (%.0f%%)" % (nm, v, f, 100 * f / v))
    rest = vals[passes * n:]
    print("  head convs:", " ".join("%.0f" % v for v in rest))
