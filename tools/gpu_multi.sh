#!/bin/bash
# Multi-GPU check: `gpurun --gpus N -- bash tools/gpu_multi.sh N [steps]`
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
N=${1:-2}; STEPS=${2:-20}
mkdir -p gpurun_out
python -m draco_b200.build > gpurun_out/env_multi.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
if [ "${SKIP_TEST:-0}" != "1" ]; then
  NCCL_DEBUG=WARN timeout -k 10 600 python -m pytest tests/test_fused_engine_gpu.py -q -m multigpu -p no:cacheprovider > gpurun_out/t_multigpu.log 2>&1; echo "multigpu test rc=$?"
fi
for impl in ours nccl; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps $STEPS --warmup 5 --impl $impl > gpurun_out/bench${N}_${impl}.log 2>&1; echo "bench $impl N=$N rc=$?"
done
tail -n 30 gpurun_out/t_multigpu.log; for impl in ours nccl; do tail -n 4 gpurun_out/bench${N}_${impl}.log; done
