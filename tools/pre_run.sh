#!/bin/bash
# Reference: tools/pre_run.sh (conda deps + hdmedians build).  Here: verify the toolchain and build the in-tree natives.
set -e
python -c "import torch, numpy; print('torch', torch.__version__, 'cuda', torch.version.cuda)"
nvcc --version | tail -1
python -m draco_b200.build --verbose
python -c "import __graft_entry__ as g; g.build(); print('natives ok')"
