#!/bin/bash
# Reference: src/data_prepare.sh -> datasets/data_prepare.py (pre-download MNIST / CIFAR-10 before the job starts).
python -m draco_b200.data.prepare "$@"
