"""Condense the JSON line of bench.py (stdin) to one readable line."""
import json, sys
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("gpus %d  Mpaths/s %.1f  ms/step %.1f  stages %s  e2e %.1f  clocks %s  launches %s" % (
    d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], {k: round(v, 1) for k, v in d["roofline"]["stage_ms_per_step"].items()},
    d["e2e"]["value"] / 1e6, d["clocks"], d["gpu_launches"]))
