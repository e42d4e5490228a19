cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/val1
python -m draco_b200.build > gpurun_out/val1/env.log 2>&1
timeout -k 10 600 python bench.py --gpus 1 --steps 5 --warmup 3 --network ResNet50 --group-size 5 --worker-fail 2 --batch-size 64 --sanity-steps 0 > gpurun_out/val1/r50.log 2> gpurun_out/val1/r50.err; echo "r50 rc=$?"
timeout -k 10 900 python bench.py --gpus 1 --steps 3 --warmup 3 --network ResNet50 --dataset ImageNet --group-size 5 --worker-fail 2 --batch-size 32 --sanity-steps 0 > gpurun_out/val1/r50_imagenet.log 2> gpurun_out/val1/r50_imagenet.err; echo "r50_imagenet rc=$?"
timeout -k 10 600 python bench.py --gpus 1 --steps 5 --warmup 3 --network VGG11 --approach cyclic --worker-fail 1 --err-mode constant --sanity-steps 0 > gpurun_out/val1/vgg_cyclic.log 2> gpurun_out/val1/vgg_cyclic.err; echo "vgg rc=$?"
tail -n 2 gpurun_out/val1/*.log | cut -c1-600
tail -n 5 gpurun_out/val1/*.err | cut -c1-300
