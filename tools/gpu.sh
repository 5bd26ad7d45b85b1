#!/bin/bash
# Rebuild everything, then run a command on the B200 box.  usage: tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
