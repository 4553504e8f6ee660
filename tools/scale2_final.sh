# 2-GPU check of the final code: captured overlapped exchange == plain schedule, then the bench at N = 2
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra-configs 2> gpurun_out/scale2.err | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*'
python bench.py --steps 20 --warmup 3 --no-extra-configs --no-cpu-baseline --no-gpu-baseline | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*'
