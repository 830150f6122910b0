"""profiles/r2_dram_traffic.json from an ncu CSV of the tcgen05 launches of ONE configs[1] step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'conv_gemm|conv3x3_halo|conv1_line' -s 53 -c 65 --csv --log-file gpurun_out/traffic.csv \
        python bench.py --quick --steps 1 --warmup 0
(-s 53 skips the 53 convolution launches of the weight-rounding calibration pass at create time; the next 65 launches are
the 53 ResNet50 layers over 2048 images + the 12 PhaseNet launches of the step.)
bench.py reads the JSON for `roofline.traffic`."""
import collections
import csv
import json
import re
import sys

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
    d = per.setdefault(row["ID"], {"kernel": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mimamo::", "")})
    d[row["Metric Name"]] = v * scale
launches = list(per.values())
total = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
images = 2048
json.dump({"dram_bytes_per_image": total / images, "launches": len(launches), "images": images, "total_dram_bytes": total,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the %d tcgen05 launches of one 2048-window step of this build "
                     "(profiles/r2_dram_traffic.json, tools_traffic.py)" % len(launches),
           "per_launch": [{"kernel": l["kernel"], "read": l.get("dram__bytes_read.sum"), "write": l.get("dram__bytes_write.sum"),
                           "us": l.get("gpu__time_duration.sum")} for l in launches]}, open(out, "w"), indent=1)
print("%d launches, %.2f GB of DRAM traffic, %.1f MB per image" % (len(launches), total / 1e9, total / images / 1e6))
