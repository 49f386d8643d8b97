import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith("{"): 
        if line: print(line[:300])
        continue
    d=json.loads(line)
    r=d.get("roofline",{})
    print(d["config"]["family"], d["config"]["logits"], "ms/step %.4f"%d["ms_per_step"], "stats_ms %.4f"%r.get("ms_per_launch",0), "GB/s %.0f frac %.3f"%(r.get("achieved",0), r.get("frac",0)), "accept %.3f"%d.get("mean_accept_length",0), "img/s %.1f"%d["value"], "e2e", d.get("e2e",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"), "lazy", d.get("lazy_stats",{}).get("ms_per_step"), d.get("clocks",{}).get("reasons"))
