"""A/B timing of the cfg3 recurrence launches (T=1000, S=16, C=320, folded form, both directions): prints median-of-5 ms
for forward and backward.  Library / variant is chosen by the environment (ASLP_B200_CUDA_LIB, ASLP_LSTM_NO_PRESPLIT)."""
import json, os, sys
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0] + "/tools")
import perf_probe as PP
tag = sys.argv[1] if len(sys.argv) > 1 else ""
f = PP.probe_lstm(1000, 16, 320, 0, 2, False)
b = PP.probe_lstm(1000, 16, 320, 0, 2, True)
print(json.dumps({"variant": tag, "fwd_us_per_step": round(f["us_per_step"], 3), "bwd_us_per_step": round(b["us_per_step"], 3)}), flush=True)
