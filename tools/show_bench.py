import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
r=d["roofline"]; print(r["kernel"], r["achieved"], r["frac"], r["conv_ms_by_family"], r["conv_total_ms"])
