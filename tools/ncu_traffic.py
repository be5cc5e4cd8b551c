"""profiles/ncu_traffic.json from the ncu --set full summaries (tools/ncu_summary.py output): dram__bytes_read.sum +
dram__bytes_write.sum per launch for each captured kernel; bench.py reads roofline.traffic from it."""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
out = {}
for path in sys.argv[1:]:
    cur = None
    for line in open(path):
        m = re.match(r"(kernel|duration|dram read|dram write)\s+(.*)", line)
        if not m:
            continue
        key, rest = m.group(1), m.group(2).strip()
        if key == "kernel":
            name = re.sub(r"<unnamed>::|void |rowreg::|recur::", "", rest).split("(")[0].split("<")[0].strip()
            base, k = name, 2
            while name in out:
                name = "%s#%d" % (base, k); k += 1
            cur = out[name] = {"source": path}
        elif cur is not None:
            val, unit = rest.split()[:2]
            if key == "duration":
                cur["ncu_duration_ms"] = float(val) * TIME[unit]
            else:
                cur["dram_%s_bytes" % key.split()[1]] = float(val) * UNIT[unit]
    for v in out.values():
        if "dram_read_bytes" in v and "dram_write_bytes" in v:
            v["dram_bytes_per_launch"] = v["dram_read_bytes"] + v["dram_write_bytes"]
json.dump(out, sys.stdout, indent=1)
