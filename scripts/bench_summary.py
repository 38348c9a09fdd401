import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.4f  cells/s %.3e  pipeline frac %.3f" % (d["ms_per_step"], d["value"], d["roofline"]["pipeline"]["frac_of_aggregate_peak"]))
print("stage_ms", d["roofline"]["stage_ms"])
if d.get("e2e"): print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
