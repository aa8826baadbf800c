timeout 300 python -m pytest tests/test_zzz_gpu_narrow_tiles.py -m gpu -x -q 2>&1 | tail -3
AB_C3_CFGS="SPIM_NOP=2|SPIM_COLP=2" bash profiles/r2_ab.sh r2_ab_c15 "SPIM_NOP=2"
