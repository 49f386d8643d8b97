#!/bin/bash
# A/B of row-statistics kernel variants built as separate libraries (lantern_b200/liblantern_b200_exp<V>.so,
# -DLANTERN_EXP=<V>); bench lines only.  usage: ab_variants.sh "0 1 2" [extra bench args]
mkdir -p gpurun_out
VARS=${1:-"0"}
Q="--no-cpu --no-torch --no-e2e --no-lazy --steps 50 --warmup 5"
for rep in 1 2; do
  for fam in lumina_mgpt llamagen; do
    for v in $VARS; do
      LANTERN_B200_LIB=$PWD/lantern_b200/liblantern_b200_exp$v.so timeout 150 python bench.py $Q --family $fam 2>/dev/null | tail -1 > gpurun_out/abv_${fam}_v${v}_r${rep}.json
    done
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/abv_*.json")):
    try:
        d = json.loads(open(f).read())
        r = d["roofline"]
        print(f, "ms_per_step=%.4f" % d["ms_per_step"], "stats_us=%.2f" % (r["ms_per_launch"] * 1e3), "frac=%.3f" % r["frac"], "accept=%.4f" % d["mean_accept_length"])
    except Exception as e:
        print(f, "ERR", e)
PY
