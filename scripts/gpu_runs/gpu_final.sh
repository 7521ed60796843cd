#!/bin/bash
# Final check of a round: smoke(), the whole GPU suite, the default bench line.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 500 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/final_suite.log 2>&1
echo "suite exit $?"; tail -n 6 gpurun_out/final_suite.log
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("BENCH", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["cpu_baseline"]["value"], d["clocks"])
PY
exit 0
