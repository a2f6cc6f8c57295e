#!/bin/bash
# round-1 final capture on one B200: tests, the bench line, per-operator benches, launch lists and ncu --set full of the operator kernels
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > $O/r1b_bench_1gpu.json 2> $O/r1b_bench_1gpu.err; echo "bench rc=$?"
python bench_ops.py --reps 3 > $O/r1b_bench_ops.jsonl 2> $O/r1b_bench_ops.err; echo "ops rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1b_launches_bench_1gpu.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu1 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1b_launches_bench_ops.csv python bench_ops.py --reps 2 --particles 2e7 > /dev/null 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fp_linear_reg<4>|k_props_reg<4>" -c 2 -o $O/r1b_full_c5 python bench_ops.py --reps 1 --particles 2e7 --only c5 > /dev/null 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge_warp|k_seg_small|k_gen_gather|k_gen_scatter_idx|k_gen_sort_segments_warp|k_gen_classify|k_ntc<" --launch-skip 0 -c 14 -o $O/r1b_full_ops python bench_ops.py --reps 2 --particles 2e7 --only ops > /dev/null 2>&1; echo "ncu4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge\(|k_ntc_warp|k_sample_on_grid" -c 4 -o $O/r1b_full_c2 python bench_ops.py --particles 2e7 --only c2 > /dev/null 2>&1; echo "ncu5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ntc<|k_convect_band|k_band_scatter|k_band_combine" --launch-skip 16 -c 4 -o $O/r1b_full_c3 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu6 rc=$?"
ls -la $O | tail -12
cut -c1-400 $O/r1b_bench_1gpu.json
