#!/bin/bash
# Run on the parameter-server node: install the cluster key, make every node reachable by alias, and replicate the
# repository (sources only; each node builds its own sm_100a extensions) to all workers.
# (role of the reference's tools/remote_script.sh:1-18)
set -euo pipefail
KEY=${1:?usage: remote_script.sh <ssh-key.pem> [user] [repo-dir]}
USER_=${2:-ubuntu}
REPO=${3:-draco_b200}
mkdir -p ~/.ssh && cp "$KEY" ~/.ssh/id_cluster && chmod 600 ~/.ssh/id_cluster
cat "$(dirname "$0")/ssh_config.cluster" >> ~/.ssh/config 2>/dev/null || true
[ -f ~/hosts ] && sudo sh -c "grep -v deeplearning-worker /etc/hosts > /tmp/hosts.new; cat /tmp/hosts.new $HOME/hosts > /etc/hosts" || true
SSH_OPTS="-o StrictHostKeyChecking=no -o UserKnownHostsFile=/dev/null -i $HOME/.ssh/id_cluster"
tail -n +2 ~/hosts_address | while read -r ip; do
  [ -z "$ip" ] && continue
  rsync -az -e "ssh $SSH_OPTS" --exclude .git --exclude gpurun_out --exclude '*.so' ~/"$REPO"/ "$USER_@$ip:~/$REPO/" &
done
wait
# build the native code everywhere (PS included)
while read -r ip; do
  [ -z "$ip" ] && continue
  ssh $SSH_OPTS "$USER_@$ip" "cd ~/$REPO && python -m draco_b200.build" < /dev/null &
done < ~/hosts_address
wait
echo "cluster ready: $(wc -l < ~/hosts_address) nodes"
