# quick perf loop: new-kernel tests, border kernel alone, bench line (no baselines), kernel split
TAG=${1:-q}
python -m pytest tests/test_bench_scale_gpu.py -m gpu -q -x -k "reflect or dot or fused" 2>&1 | tail -2
python tools/bench_border.py
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extra-configs > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python tools/profile_torch.py --top 16 2>/dev/null > gpurun_out/${TAG}_prof.txt; grep -i "border\|total GPU" gpurun_out/${TAG}_prof.txt
