#!/bin/bash
# Developer tool: leave a broken GPU box quickly instead of hanging in the first CUDA call.
timeout 90 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader || { echo "GPU box unhealthy (nvidia-smi)"; exit 9; }
timeout 120 python -c "import torch; torch.zeros(4, device='cuda').sum().item(); print('cuda ok')" || { echo "GPU box unhealthy (first CUDA call)"; exit 9; }
