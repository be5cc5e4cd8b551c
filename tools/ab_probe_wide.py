"""A/B timing of the wide backward recurrence (8*NS streams per chain): cfg2-like (100 streams, 512 cells, one direction)
and a 32-stream bidirectional 320-cell case; T = 200 so the launch is not the measurement."""
import json, os, sys
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0] + "/tools")
import perf_probe as PP
tag = sys.argv[1] if len(sys.argv) > 1 else ""
a = PP.probe_lstm(200, 100, 512, 0, 1, True)
b = PP.probe_lstm(200, 32, 320, 0, 2, True)
c = PP.probe_lstm(200, 24, 320, 0, 2, True)
print(json.dumps({"variant": tag, "bwd_us_per_step_S100_C512": round(a["us_per_step"], 3), "bwd_us_per_step_S32_C320x2": round(b["us_per_step"], 3),
                  "bwd_us_per_step_S24_C320x2": round(c["us_per_step"], 3)}), flush=True)
