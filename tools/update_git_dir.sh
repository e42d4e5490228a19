#!/bin/bash
# Bring every node of the cluster to the PS node's checkout (role of the reference's tools/update_git_dir.sh:1-10):
# rsync the working tree, then rebuild the in-tree extensions on each node.
set -euo pipefail
REPO=${1:-draco_b200}
USER_=${2:-ubuntu}
SSH_OPTS="-o StrictHostKeyChecking=no -o UserKnownHostsFile=/dev/null"
tail -n +2 ~/hosts_address | while read -r ip; do
  [ -z "$ip" ] && continue
  ( rsync -az -e "ssh $SSH_OPTS" --exclude .git --exclude gpurun_out --exclude '*.so' ~/"$REPO"/ "$USER_@$ip:~/$REPO/" &&
    ssh $SSH_OPTS "$USER_@$ip" "cd ~/$REPO && python -m draco_b200.build" < /dev/null ) &
done
wait
