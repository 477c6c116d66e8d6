"""one-screen summary of a bench.py line: python tools/bench_brief.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f reads/s  step %.1f ms  stages %s" % (d["value"], d["ms_per_step"], json.dumps({k: round(v, 1) for k, v in d["stages_ms"].items()})))
print("e2e %.0f reads/s  %s" % (d["e2e"]["value"], json.dumps({k: round(v, 1) for k, v in d["e2e"]["stages_ms"].items()})))
print(json.dumps(d["parity_vs_planted"]), "poa GCUPS %.1f" % d["kernels"]["k_poa"]["GCUPS"], "roofline:", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
