set -x
# sweep of BASELINE configs[4] with the rewritten lazy walk
python profiles/sweep_accept.py --out gpurun_out/sweep_r2b.json > gpurun_out/sweep_r2b.log 2>&1
# launch list of the timed region (default automatic schedule: one lazy walk kernel per step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b_auto.csv python bench.py --no-cpu --no-torch --no-e2e --no-extra --no-lazy --steps 2 --warmup 3 > /dev/null 2>&1
# full capture of the lazy walk kernel (the kernel the headline times); summarised on the box, the report itself is too
# large to travel back
ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 4 -c 1 -f -o /tmp/r2b_walk_lazy python bench.py --no-cpu --no-torch --no-e2e --no-extra --no-lazy --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r2b_walk_lazy.ncu-rep > gpurun_out/r2b_walk_lazy_summary.txt 2>&1
python profiles/ncu_lines.py /tmp/r2b_walk_lazy.ncu-rep 40 >> gpurun_out/r2b_walk_lazy_summary.txt 2>&1
python profiles/dropin_latency.py > gpurun_out/dropin_latency_r2b.json 2> gpurun_out/dropin_latency_r2b.err
ls -la gpurun_out | tail -12
