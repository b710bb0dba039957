"""Per-layer conv throughput INSIDE the sustained (power-capped) benchmark step: run bench.py with
STORM_B200_PROFILE_DUMP=<csv>, then: python tools/layer_sustained.py <csv>  (launch order repeats every forward)."""
import collections
import sys

rows = [l.strip().split(",") for l in open(sys.argv[1]) if l.strip()]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 49
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for i, (eng, ms, fl) in enumerate(rows):
    a = agg[i % per]
    a[0] += float(ms); a[1] += float(fl); a[2] += 1
tot_ms = sum(a[0] for a in agg.values())
print(f"{len(rows)} conv launches, {per} per forward, total {tot_ms:.1f} ms")
for k in sorted(agg):
    ms, fl, n = agg[k]
    print(f"conv #{k:2d}: {1e3 * ms / n:8.1f} us  {fl / n / 1e9:8.1f} GFLOP  {fl / ms / 1e9:7.1f} TFLOP/s alg ({3 * fl / ms / 1e9:7.1f} issued)  {100 * ms / tot_ms:5.1f}%")
