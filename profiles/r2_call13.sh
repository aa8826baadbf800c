AB_C3_CFGS="SPIM_NOP=2|SPIM_COL_NARROW=1" bash profiles/r2_ab.sh r2_ab_c13
