#!/bin/bash
# Optional operator tooling: pdsh for fan-out shell commands (role of the reference's tools/install.sh:1-14).
# Needs network access; everything the framework itself needs is already in the image.
set -euo pipefail
if command -v pdsh > /dev/null; then echo "pdsh already installed"; exit 0; fi
sudo apt-get update && sudo apt-get install -y pdsh
echo "export PDSH_RCMD_TYPE=ssh" >> ~/.bashrc
