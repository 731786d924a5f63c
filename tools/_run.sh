set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mrhs or degenerate" 2>&1 | tail -3
for cfg in 64 32 16; do
  HIFIR_B200_MRHS_WIDE=$cfg timeout 200 python tools/mrhs_bench.py 2>&1 | grep -v "^\[bench\]" | tee -a gpurun_out/mrhs_cols_3.log
done
