AB_C3_CFGS="SPIM_NOP=2|SPIM_NOP=3" bash profiles/r2_ab.sh r2_ab_c17 "SPIM_NOP=2"
