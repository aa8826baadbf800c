timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c14.txt 2>&1; tail -3 gpurun_out/r2_gputests_c14.txt
AB_C3_CFGS="SPIM_NOP=2|SPIM_XFWD_LINES=16|SPIM_XFWD_TN=192|SPIM_XFWD_TN=160" bash profiles/r2_ab.sh r2_ab_c14 "SPIM_NOP=2" "SPIM_XFWD_LINES=8"
