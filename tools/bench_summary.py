import json, sys
d = json.load(sys.stdin)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
print(tag, "value", round(d["value"] / 1e6, 2), "M/s  ms/step", round(d["ms_per_step"], 4), " e2e", round(d["e2e"]["value"] / 1e6, 2),
      {k: round(v * 1e3, 1) for k, v in d["roofline"]["kernels_ms"].items()}, "launches", d["gpu_launches"])
