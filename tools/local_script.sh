#!/bin/bash
# Run on the operator machine: push the cluster key and the host lists written by `python -m draco_b200.cli.cluster get_hosts`
# to the parameter-server node, then hand over to remote_script.sh there.
# (role of the reference's tools/local_script.sh:1-10; key path, user and repo dir are arguments here, not hard-coded)
set -euo pipefail
KEY=${1:?usage: local_script.sh <ssh-key.pem> [user] [repo-dir-on-node]}
USER_=${2:-ubuntu}
REMOTE_DIR=${3:-draco_b200}
PS_HOST=$(head -n 1 hosts_address)
SSH_OPTS="-o StrictHostKeyChecking=no -o UserKnownHostsFile=/dev/null -i $KEY"
scp $SSH_OPTS "$KEY" hosts hosts_alias hosts_address "$USER_@$PS_HOST:~/"
rsync -az -e "ssh $SSH_OPTS" --exclude .git --exclude gpurun_out --exclude '*.so' "$(dirname "$0")/.." "$USER_@$PS_HOST:~/$REMOTE_DIR"
ssh $SSH_OPTS "$USER_@$PS_HOST" "bash ~/$REMOTE_DIR/tools/remote_script.sh ~/$(basename "$KEY") $USER_ $REMOTE_DIR"
