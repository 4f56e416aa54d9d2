#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 12
python __graft_entry__.py smoke 2>&1 | tail -n 2
