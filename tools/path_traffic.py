"""DRAM traffic of ONE whole NCSN++ evaluation from an ncu launch list taken with
   ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
       --csv --log-file gpurun_out/traffic.csv python tools/profile_forward.py          (B = 8, 256x256, ncsnpp)
Usage: python tools/path_traffic.py gpurun_out/traffic.csv profiles/r2_path_traffic.json
Writes the JSON bench.py reads for roofline.traffic (dominant kernel, per launch) and the whole path's bytes per utterance per
evaluation, next to the algorithmic 0.879 GB (ideal fusion) / 1.793 GB (two-kernel split) of SURVEY.md section 8(d)."""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    kid = row["ID"]
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"].lower()
    m = row["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    else:
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    per.setdefault(kid, dict(name=name))[m] = v
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in per.values():
    a = by[k["name"]]
    a[0] += 1
    a[1] += k.get("dram__bytes_read.sum", 0.0)
    a[2] += k.get("dram__bytes_write.sum", 0.0)
    a[3] += k.get("gpu__time_duration.sum", 0.0)
tot_b = sum(a[1] + a[2] for a in by.values())
tot_us = sum(a[3] for a in by.values())
conv = {n: a for n, a in by.items() if "conv_tc" in n}
conv_n = sum(a[0] for a in conv.values())
conv_b = sum(a[1] + a[2] for a in conv.values())
out = dict(
    batch=B, launches=len(per), path_dram_bytes_per_eval=tot_b, path_dram_bytes_per_utt_nfe=tot_b / B, path_us_serialised=tot_us,
    algorithmic_bytes_per_utt_nfe=dict(ideal_fused=0.879e9, two_kernel_split=1.793e9),
    conv_launches=conv_n, conv_dram_bytes_per_launch=conv_b / max(1, conv_n), conv_dram_bytes_per_eval=conv_b,
    by_kernel={n: dict(launches=a[0], read_bytes=a[1], write_bytes=a[2], us=a[3]) for n, a in sorted(by.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))},
    note=("ncu dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one NCSN++ evaluation (B=8, 256x256, ncsnpp); "
          "`conv_dram_bytes_per_launch` = mean over the tcgen05 conv launches of that evaluation (the kernel class `roofline` is about)"))
json.dump(out, open(dst, "w"), indent=1)
print(f"{len(per)} launches, {tot_us:.0f} us serialised, path DRAM {tot_b / 1e9:.2f} GB per evaluation = {tot_b / B / 1e9:.3f} GB per utterance "
      f"(algorithmic 0.879 ideal / 1.793 two-kernel); convs {conv_b / 1e9:.2f} GB over {conv_n} launches")
for n, a in sorted(by.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))[:12]:
    print(f"  {(a[1] + a[2]) / 1e6:9.1f} MB  {a[3]:9.1f} us  n={a[0]:3d}  {n}")
