# 2-GPU bench runs under different NCCL CTA budgets (gpurun --gpus 2 -- bash tools/scale2.sh)
for cfg in "default" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=1"; do
  echo "== $cfg"
  if [ "$cfg" = "default" ]; then E=""; else E="$cfg"; fi
  env $E timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra-configs 2> gpurun_out/scale2.err | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*'
done
echo "== 1 GPU"
python bench.py --steps 20 --warmup 3 --no-extra-configs --no-cpu-baseline --no-gpu-baseline | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*'
