timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "solve or krylov or fgmres or hifir or apply" 2>&1 | tail -3
for g in 1 0; do HIFIR_B200_GRAPH=$g timeout 200 python tools/tune_sweep.py --size 128 --steps 100 --cfg "" 2>&1 | grep -A1 "apply\|failed" | sed "s/^/[graph=$g] /"; done
