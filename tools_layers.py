"""Label the per-launch GEMM times printed by `bench.py --layers` (stderr) with ResNet50 layer names and floors."""
import sys
TF, BW = 1363.5e12, 6557.8e9
B = 512
names, floors = ["conv1"], [max(2 * B * 12544 * 147 * 64 / TF, (B * 115 * 115 * 16 + B * 12544 * 64) * 2 / BW) * 1e6]
cin = 64
for si, (nb, mid, H) in enumerate([(3, 64, 56), (4, 128, 28), (6, 256, 14), (3, 512, 7)]):
    cout = mid * 4
    for b in range(nb):
        M = B * H * H
        def add(n, K, N, rd, wr):
            names.append("s%db%d %s" % (si + 2, b + 1, n)); floors.append(max(2 * M * K * N / TF, (rd + wr) * 2 / BW) * 1e6)
        add("reduce", cin, mid, M * cin, M * mid)
        add("3x3", 9 * mid, mid, M * mid, M * mid)
        if b == 0: add("proj", cin, cout, M * cin, M * cout)
        add("incr", mid, cout, M * mid + M * cout, M * cout)
        cin = cout
for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.strip() and l.strip()[0].isdigit()]
    vals = [float(x) for x in lines[-1].split()]
    n = len(names)
    passes = max(1, (len(vals) - 12) // n)                       # ResNet passes per step (2048 / images per pass)
    B = 2048 // passes
    chunk = [sum(vals[c * n + i] for c in range(passes)) / passes * (512.0 / B) for i in range(n)]      # per 512 images
    print("== %s: %d launches per step, %d-image ResNet passes; per 512 images: %.0f us measured, %.0f us floor (sum of per-layer max(tensor, HBM))" % (path, len(vals), B, sum(chunk), sum(floors)))
    for nm, v, f in zip(names, chunk, floors):
        print("  %-12s %7.1f us  floor %6.1f  (%.0f%%)" % (nm, v, f, 100 * f / v))
    rest = vals[passes * n:]
    print("  head convs:", " ".join("%.0f" % v for v in rest))
