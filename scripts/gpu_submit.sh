#!/bin/bash
# Submit a GPU job from a FROZEN copy of the tree: gpurun snapshots /root/repo only when a box is granted (after the queue),
# so edits made while a call waits would leak into it.  The copy under .frozen/<id>/ is taken now; the remote command runs there.
#   scripts/gpu_submit.sh <timeout_s> <script.sh> [gpus]
set -e
cd "$(dirname "$0")/.."
id=$(date +%H%M%S)
rm -rf .frozen
mkdir -p .frozen/$id
tar --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude=__pycache__ --exclude=./.pytest_cache -cf - . | tar -xf - -C .frozen/$id
gpus=${3:-1}
exec /usr/local/graft/bin/gpurun --gpus $gpus --timeout $1 -- "mkdir -p gpurun_out && cd .frozen/$id && ln -sfn ../../gpurun_out gpurun_out && python -m satlas_super_resolution_b200.build > /dev/null && bash $2"
