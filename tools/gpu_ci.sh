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
    gemmbench) timeout -k 10 600 python tools/bench_gemm.py --json gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1; echo "gemmbench rc=$?" ;;
    ncu_gemm) timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 2 -f -o gpurun_out/prof_gemm python tools/bench_gemm.py --quick > gpurun_out/ncu_gemm.log 2>&1; echo "ncu_gemm rc=$?" ;;
    ncu_step) timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"push_encode|vote_compare|aggregate_update|cast_params" -s 10 -c 10 -f -o gpurun_out/prof_step python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu_step rc=$?" ;;
    ncu_bn) timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"bn_stats|bn_apply|bn_bwd" -s 80 -c 8 -f -o gpurun_out/prof_bn python tools/prof_step.py > gpurun_out/ncu_bn.log 2>&1; echo "ncu_bn rc=$?" ;;
    ncu_cyclic) timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"cyclic_project|cyclic_locate|aggregate_update|push_encode" -s 9 -c 10 -f -o gpurun_out/prof_cyclic python tools/prof_step.py cyclic > gpurun_out/ncu_cyclic.log 2>&1; echo "ncu_cyclic rc=$?" ;;
    launches) STEPS=3 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1; echo "launches rc=$?" ;;
    worker) timeout -k 10 600 python tools/bench_worker.py > gpurun_out/bench_worker.log 2>&1; echo "worker rc=$?" ;;
    diag) timeout -k 10 120 python tools/diag_flags.py > gpurun_out/diag_flags.log 2>&1; echo "diag rc=$?" ;;
    norm) timeout -k 10 300 python -m pytest tests/test_norm_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_norm.log 2>&1; echo "norm rc=$?" ;;
    convbench) timeout -k 10 420 python tools/bench_conv.py > gpurun_out/conv_bench.log 2>&1; echo "convbench rc=$?" ;;
    conv) timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py tests/test_norm_gpu.py -q -m gpu -p no:cacheprovider -k "conv or bn or norm" > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" ;;
    ncu_conv) timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"convg|conv_halo|wgradg|stem_|bn_" -s 0 -c 60 -f -o gpurun_out/prof_conv python tools/prof_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu_conv rc=$?" ;;
    sanitizer) for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do timeout -k 10 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "push or vote or aggregate or geometric or krum or cyclic" > gpurun_out/sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; done ;;
    geomed) timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py tests/test_fused_engine_gpu.py -q -m gpu -p no:cacheprovider -k "geomed or krum or geometric" > gpurun_out/t_geomed.log 2>&1; echo "geomed rc=$?"
            timeout -k 10 300 python bench.py --gpus 1 --steps 30 --warmup 5 --approach baseline --mode geometric_median > gpurun_out/bench1_geomed.log 2>&1; echo "bench geomed rc=$?" ;;
    alltests) timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "alltests rc=$?" ;;
  esac
done
tail -n 25 gpurun_out/t_*.log gpurun_out/bench*.log gpurun_out/diag*.log 2>/dev/null | cut -c1-1200 | tail -n 150
