"""print a compact summary of the last bench.py JSON line read from stdin"""
import json, sys
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line); r = d["roofline"]
print(sys.argv[1] if len(sys.argv) > 1 else "", d["config"]["reads_per_gpu"],
      "value %.0f reads/s  kernel_ms %.1f  e2e %.0f  blocks %.3g text %.3g ext %.3g frac %.3f" %
      (d["value"], r["kernel_ms"], d["e2e"]["value"], r.get("index_blocks", 0), r.get("text_extensions", 0),
       d["config"]["extensions_per_step_rank0"], r["frac"]))
