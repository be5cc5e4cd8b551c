"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share for the LAST
step of a `bench.py --steps 1 --warmup W` run (launches are split evenly over the W+1 steps)."""
import collections
import csv
import sys

fn, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
lines = [l for l in open(fn) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
n = len(rows) // nsteps
# the training steps are identical: their period is the spacing of a kernel that runs once per step (the CTC label CSR
# build, or the one-launch CTC kernel); the summary covers the last whole period of the run (profiling runs use --no-secondary, nothing follows it)
marks = [i for i, x in enumerate(rows) if "ctc_csr_kernel" in x["Kernel Name"] or "ctc_fused_kernel" in x["Kernel Name"]]
if len(marks) >= 2:
    n = marks[-1] - marks[-2]
    last = rows[marks[-2] + 1:marks[-1] + 1]           # one whole period: from behind one CTC launch up to and including the next
else:
    last = rows[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for x in last:
    k = x["Kernel Name"]
    k = k.replace("<unnamed>::", "").replace("void ", "")
    k = k.split("(")[0][:70]
    v = float(x["Metric Value"]); u = x["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("launches in step: %d   sum of kernel durations: %.3f ms (cold-cache, serialised: use shares, not absolutes)" % (n, tot))
print("%-72s %6s %10s %7s" % ("kernel", "n", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %10.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
