set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2b_tests.log; cat gpurun_out/r2b_tests.log
( for tool in memcheck racecheck synccheck initcheck; do echo "== compute-sanitizer --tool $tool python profiles/sanitize_cases.py"; timeout 1200 compute-sanitizer --tool $tool python profiles/sanitize_cases.py 2>&1 | grep -v "^=========     \(Saved\|at \|by \|in \)" | cut -c1-260 | tail -40; done ) > gpurun_out/r2b_sanitizer.txt 2>&1
grep -c "lazy_probs\|walk_kernel\|lazy_lift" gpurun_out/r2b_sanitizer.txt; grep "SUMMARY" gpurun_out/r2b_sanitizer.txt
python -c "import __graft_entry__ as g; g.smoke()"
