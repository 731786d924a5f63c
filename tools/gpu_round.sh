#!/bin/bash
# Developer script for one gpurun call: parity subset, then A/B timings of attach-time configurations.
# Every step runs under its own timeout: a hung kernel must never eat the GPU budget.
set -u
mkdir -p gpurun_out
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest subset"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${PYTEST_K:-solve}" 2>&1 | tail -15
echo "== tune"
for cfg in "$@"; do
  timeout ${CFG_TIMEOUT:-150} python tools/tune_sweep.py --size ${SIZE:-128} --cfg "$cfg" 2>&1 | grep -v "^\[bench\] reference" | tee -a gpurun_out/tune_$(date +%H%M).log
  rc=${PIPESTATUS[0]}; [ $rc -ne 0 ] && echo "[$cfg] rc=$rc (124 = timeout)"
done
