# 8-GPU bench runs: overlapped vs serial gradient exchange (gpurun --gpus 8 -- bash tools/scale8.sh)
for cfg in "OTM_DDP_OVERLAP=1" "OTM_DDP_OVERLAP=0" "OTM_DDP_OVERLAP=0 NCCL_ALGO=Tree"; do
  echo "== $cfg"
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra-configs 2> gpurun_out/scale8.err | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*'
done
