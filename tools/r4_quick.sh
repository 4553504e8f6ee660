# quick perf loop: reflect/dot tests, bench line (no baselines), kernel splits at 128 and 256
TAG=${1:-q}
python -m pytest tests/test_bench_scale_gpu.py -m gpu -q -x -k "reflect or dot or fused" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [c["value"] for c in d.get("configs",[])], d["roofline"]["frac"])
PY
python tools/bench_border.py; python tools/profile_torch.py --top 16 2>/dev/null > gpurun_out/${TAG}_prof.txt; head -18 gpurun_out/${TAG}_prof.txt
python tools/profile_torch.py --top 16 --size 256 2>/dev/null > gpurun_out/${TAG}_prof256.txt; head -18 gpurun_out/${TAG}_prof256.txt
