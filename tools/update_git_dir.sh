#!/bin/bash
# Bring every node of the cluster to the PS node's checkout (role of the reference's tools/update_git_dir.sh:1-10):
# rsync the working tree, then rebuild the in-tree extensions on each node.  Uses the node aliases + the dedicated ssh config
# installed by remote_script.sh.
set -euo pipefail
REPO=${1:-draco_b200}
USER_=${2:-ubuntu}
SSH="ssh -F $HOME/.ssh/config.draco_cluster"
tail -n +2 ~/hosts_alias | while read -r node; do
  [ -z "$node" ] && continue
  ( rsync -az -e "$SSH" --exclude .git --exclude gpurun_out --exclude '*.so' ~/"$REPO"/ "$USER_@$node:~/$REPO/" &&
    $SSH "$USER_@$node" "cd ~/$REPO && python -m draco_b200.build" < /dev/null ) &
done
wait
