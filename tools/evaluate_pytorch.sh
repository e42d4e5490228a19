#!/bin/bash
# Reference: src/evaluate_pytorch.sh -- follow a run through its checkpoints.
python -m draco_b200.cli.distributed_evaluator --eval-batch-size=1000 --eval-freq=${EVAL_FREQ:-200} --network=${NETWORK:-FC} \
  --dataset=${DATASET:-MNIST} --model-dir=output/models/ "$@"
