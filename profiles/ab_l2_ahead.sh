#!/bin/bash
# A/B of the row-statistics kernel's L2 look-ahead (LANTERN_STATS_L2_AHEAD) on one B200; bench lines only.
mkdir -p gpurun_out
Q="--no-cpu --no-torch --no-e2e --no-lazy --steps 50 --warmup 5"
for rep in 1 2; do
  for fam in lumina_mgpt llamagen; do
    for a in 0 1; do
      LANTERN_STATS_L2_AHEAD=$a timeout 150 python bench.py $Q --family $fam 2>/dev/null | tail -1 > gpurun_out/ab_${fam}_a${a}_r${rep}.json
    done
  done
done
LANTERN_STATS_L2_AHEAD=1 timeout 150 python bench.py $Q --items 256 2>/dev/null | tail -1 > gpurun_out/ab_lumina256_a1.json
LANTERN_STATS_L2_AHEAD=0 timeout 150 python bench.py $Q --items 256 2>/dev/null | tail -1 > gpurun_out/ab_lumina256_a0.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read())
        r = d["roofline"]
        print(f, "ms_per_step=%.4f" % d["ms_per_step"], "stats_us=%.2f" % (r["ms_per_launch"] * 1e3), "frac=%.3f" % r["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
