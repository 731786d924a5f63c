for g in 1 0; do for f in 1 0; do HIFIR_B200_GRAPH=$g HIFIR_B200_MGS_FUSED=$f timeout 200 python tools/_k.py 2>/dev/null; done; done
