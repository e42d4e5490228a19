#!/bin/bash
# The reference's shipped example job (src/run_pytorch.sh: FC/MNIST, cyclic code, n=7 workers, s=2, constant adversary,
# compression on), on one 8-GPU node: 8 processes = 1 PS + 7 workers, fused NVLink transport.
NPROC=${NPROC:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node=${NPROC} --master-addr 127.0.0.1 --master-port ${PORT:-29500} \
  -m draco_b200.cli.distributed_nn \
  --lr=0.01 --momentum=0.9 --network=FC --dataset=MNIST --batch-size=4 --comm-type=Bcast --mode=normal \
  --approach=cyclic --eval-freq=200 --err-mode=constant --adversarial=1 --epochs=50 --max-steps=${MAX_STEPS:-1000} \
  --worker-fail=2 --group-size=3 --compress-grad=compress --checkpoint-step=0 --num-workers=7 \
  --train-dir=output/models/ "$@"
