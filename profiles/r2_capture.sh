set -x
# launch list of the timed region (streamed schedule: both kernels; default auto schedule: one kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_streamed.csv python bench.py --no-cpu --no-torch --no-e2e --no-extra --no-lazy --schedule streamed --steps 2 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_auto.csv python bench.py --no-cpu --no-torch --no-e2e --no-extra --no-lazy --steps 2 --warmup 3 > /dev/null 2>&1
# full capture of the dominant kernel (fp32, default workload)
ncu --set full --clock-control none --import-source on -k regex:row_stats_stream --launch-skip 6 -c 1 -f -o gpurun_out/r2_row_stats_stream python bench.py --no-cpu --no-torch --no-e2e --no-extra --no-lazy --no-graph --schedule streamed --steps 3 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out
