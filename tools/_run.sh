set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for sl in 0 100 300; do
  HIFIR_B200_MRHS_WIDE=64 HIFIR_B200_MRHS_POLL_SLEEP=$sl timeout 200 python tools/mrhs_bench.py 2>&1 | grep -v "^\[bench\]" | grep -v "columns vs" | sed "s/^/[sleep $sl] /" | tee -a gpurun_out/mrhs_cols_6.log
done
timeout 300 python tools/tune_sweep.py --size 128 --cfg "" 2>&1 | grep -v "^\[bench\] reference" | tee -a gpurun_out/tune_final_a.log
