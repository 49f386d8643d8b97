#!/bin/bash
# quick A/B of the row-statistics kernel: bench lines for the four standard shapes (fp32 / bf16 Lumina, LlamaGen, 256 prompts)
for a in "" "--logits-dtype bf16" "--family llamagen" "--items 256" "--family llamagen --logits-dtype bf16" "--items 256 --logits-dtype bf16"; do
  python bench.py --no-cpu --no-torch --no-e2e --no-lazy --steps 30 $a 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-40s step %.4f ms  stats %.4f ms  frac %.3f' % ('$a', d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']))"
done
