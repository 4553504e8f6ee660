# full GPU test suite + the default bench line (gpurun -- bash tools/full_check.sh [tag])
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("configs"), d["roofline"]["frac"], d.get("gpu_baseline",{}).get("tf32"), d.get("cpu_baseline",{}).get("value"))
PY
