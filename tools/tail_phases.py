"""Prints bench.py's tail phase table (differences between the cumulative stamps) from a bench JSON line."""
import json
import sys
for path in sys.argv[1:]:
    d = json.load(open(path))
    ph = d.get("tail_phases_us") or {}
    parts = [f"{k} {v:.1f}" for k, v in ph.items()]
    print(f"{path}: step {d['ms_per_step']*1e3:.1f} us, sensor {d['stage_ms'].get('sensor', 0)*1e3:.1f}, "
          f"tail {d['stage_ms'].get('tail', 0)*1e3:.1f} | " + ", ".join(parts))
