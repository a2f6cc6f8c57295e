#!/bin/bash
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fp_linear_reg|k_props_reg" -c 2 -o $O/r1b_full_c5 python bench_ops.py --reps 1 --particles 2e7 --only c5 > $O/ncu_c5.log 2>&1; echo "ncu3 rc=$?"; tail -2 $O/ncu_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_merge$|^k_ntc$" -c 3 -o $O/r1b_full_c2b python bench_ops.py --particles 2e7 --only c2 > $O/ncu_c2b.log 2>&1; echo "ncu5 rc=$?"; tail -2 $O/ncu_c2b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_ntc$|k_convect_band|k_band_scatter|k_band_combine" --launch-skip 12 -c 4 -o $O/r1b_full_c3 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/ncu_c3.log 2>&1; echo "ncu6 rc=$?"; tail -3 $O/ncu_c3.log
ls -la $O/*.ncu-rep
