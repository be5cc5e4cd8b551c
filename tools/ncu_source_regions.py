"""Summarise the SASS-level warp-state samples of one kernel in an `ncu --set full --import-source on` report: share of
samples before / inside / after the tensor-core contraction (delimited by the first and last HMMA), top stall reasons per
region and the instructions that collect the most samples.  Usage: ncu_source_regions.py report.ncu-rep [kernel-name-substring] > profiles/..."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# a report with several kernels lists them one after the other, each behind its own "Kernel Name" row; argv[2] picks one
want = sys.argv[2] if len(sys.argv) > 2 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > pick), len(rows))
rows = rows[pick:end]
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, "# Samples") for r in data)
hm = [i for i, r in enumerate(data) if "HMMA" in r[ix["Source"]]]
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[ix["Source"]]]
first_bar = max([b for b in bars if b < hm[0]], default=hm[0] - 1)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total warp-state samples %d over %d SASS instructions; %d HMMA between instruction %d and %d" % (tot, len(data), len(hm), hm[0], hm[-1]))


def region(a, b, name):
    s = sum(num(r, "# Samples") for r in data[a:b])
    st = sorted(((k, sum(num(r, k) for r in data[a:b])) for k in stalls), key=lambda x: -x[1])[:5]
    print("%-46s %4d instr  %5.1f %% of samples   %s" % (name, b - a, 100.0 * s / tot, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in st)))


region(0, first_bar, "loop top: exchange wait, operand loads")
region(first_bar, first_bar + 4, "barrier before the contraction")
region(first_bar + 4, hm[-1] + 1, "contraction")
region(hm[-1] + 1, len(data), "partial sums, cell chain, stores, next copies")
print("\ninstructions with the most samples:")
for i in sorted(sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:14]):
    r = data[i]
    st = sorted(((k, num(r, k)) for k in stalls), key=lambda x: -x[1])[:2]
    print("  #%-5d %5.1f %%  %-64s %s" % (i, 100.0 * num(r, "# Samples") / tot, r[ix["Source"]].strip()[:64], ", ".join("%s %d" % (k[6:], v) for k, v in st)))
