#!/bin/bash
# Multi-GPU oracle-parity tests on hardware (NCCL pack path, direct push over peer memory, one process with a host thread
# per device); the log is kept under profiles/.   gpurun --gpus N -- bash profiles/r2_multigpu.sh <tag>
set -u
tag=${1:-r2_multigpu}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/$tag.txt 2>&1
nvidia-smi topo -m >> gpurun_out/$tag.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_nccl.py -m gpu -v -rA -s >> gpurun_out/$tag.txt 2>&1
echo "exit $?" >> gpurun_out/$tag.txt
tail -15 gpurun_out/$tag.txt
