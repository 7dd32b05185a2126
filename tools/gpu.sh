#!/bin/bash
# Rebuild the library locally (the .so travels with the snapshot), then run a command on the GPU box.
#   tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
