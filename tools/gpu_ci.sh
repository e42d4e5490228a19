#!/bin/bash
# GPU check run by `gpurun -- bash tools/gpu_ci.sh [stage ...]`; every stage logs to gpurun_out/ and never blocks the next one.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
STAGES="${@:-kernels gemm engine bench}"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.csv 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.device_count())" > gpurun_out/env.log 2>&1
python -m draco_b200.build >> gpurun_out/env.log 2>&1
for s in $STAGES; do
  case $s in
    kernels) timeout -k 10 420 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" ;;
    gemm)    timeout -k 10 420 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" ;;
    engine)  timeout -k 10 900 python -m pytest tests/test_fused_engine_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_engine.log 2>&1; echo "engine rc=$?" ;;
    bench)   timeout -k 10 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?" ;;
    benchnccl) timeout -k 10 600 python bench.py --gpus 1 --steps 10 --warmup 3 --impl nccl > gpurun_out/bench1_nccl.log 2>&1; echo "benchnccl rc=$?" ;;
    alltests) timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "alltests rc=$?" ;;
  esac
done
tail -n 25 gpurun_out/t_*.log gpurun_out/bench*.log 2>/dev/null | tail -n 150
