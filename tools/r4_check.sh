# new reflect-dgrad tests first, then the full GPU suite and the default bench line
TAG=${1:-r4}
python -m pytest tests/test_bench_scale_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "reflect or dot or fused or multi" > gpurun_out/${TAG}_new.log 2>&1; tail -15 gpurun_out/${TAG}_new.log
bash tools/full_check.sh ${TAG}
python tools/profile_torch.py > gpurun_out/${TAG}_prof.txt 2>&1; head -30 gpurun_out/${TAG}_prof.txt
