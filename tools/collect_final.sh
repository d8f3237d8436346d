#!/bin/bash
# After tools/final_profiles.sh came back: summaries into profiles/ (runs here, no GPU).
set -e
O=gpurun_out
rm -f profiles/ncu_traffic.json
python tools/ncu_summary.py $O/r2f_wgrad_L2_96x48.ncu-rep "k_wgrad_tc (final round-2 state), submanifold wgrad level 2, 96->48, 102714 rows; python tools/run_conv_layer.py --kind smc --level 2 --cin 96 --cout 48 --dir wgrad" --traffic-key "conv smc wgrad 96->48 rows 102714 tf32" --src-file mm2d3d_b200/csrc/conv_tc_wgrad.cu --alg-bytes 70754040 --out-name r2f_ncu_wgrad_tc_L2_96x48.txt > profiles/r2f_ncu_wgrad_tc_L2_96x48.txt
python tools/ncu_summary.py $O/r2f_wgrad_L1_64x32.ncu-rep "k_wgrad_tc (final round-2 state), submanifold wgrad level 1, 64->32, 169176 rows -- the longest launch of the step in the final bench line; python tools/run_conv_layer.py --kind smc --level 1 --cin 64 --cout 32 --dir wgrad" --traffic-key "conv smc wgrad 64->32 rows 169176 tf32" --src-file mm2d3d_b200/csrc/conv_tc_wgrad.cu --alg-bytes 83455776 --out-name r2f_ncu_wgrad_tc_L1_64x32.txt > profiles/r2f_ncu_wgrad_tc_L1_64x32.txt
python tools/ncu_summary.py $O/r2f_wgrad_rings_L0_32x16.ncu-rep "k_wgrad_tc_rings (two-ring variant for c_in <= 32), submanifold wgrad level 0, 32->16, 238021 rows; python tools/run_conv_layer.py --kind smc --level 0 --cin 32 --cout 16 --dir wgrad" --traffic-key "conv smc wgrad 32->16 rows 238021 tf32" --src-file mm2d3d_b200/csrc/conv_tc_wgrad.cu --alg-bytes 71461596 --out-name r2f_ncu_wgrad_rings_L0_32x16.txt > profiles/r2f_ncu_wgrad_rings_L0_32x16.txt
python tools/ncu_summary.py $O/r2f_conv_tc_L0_16x16.ncu-rep "k_conv_tc (final round-2 state), submanifold fwd level 0, 16->16, 238021 rows; python tools/run_conv_layer.py --kind smc --level 0 --cin 16 --cout 16 --dir fwd" --traffic-key "conv smc fwd 16->16 rows 238021 tf32" --src-file mm2d3d_b200/csrc/conv_tc.cu --alg-bytes 56200604 --out-name r2f_ncu_conv_tc_L0_16x16.txt > profiles/r2f_ncu_conv_tc_L0_16x16.txt
python tools/ncu_summary.py $O/r2f_conv_tc_L2_96x48.ncu-rep "k_conv_tc (final round-2 state), submanifold fwd level 2, 96->48, 102714 rows; python tools/run_conv_layer.py --kind smc --level 2 --cin 96 --cout 48 --dir fwd" --traffic-key "conv smc fwd 96->48 rows 102714 tf32" --src-file mm2d3d_b200/csrc/conv_tc.cu --alg-bytes 70754040 --out-name r2f_ncu_conv_tc_L2_96x48.txt > profiles/r2f_ncu_conv_tc_L2_96x48.txt
python tools/summarize_launches.py $O/r2_launches.csv > profiles/r2_tf32_launches.txt
tail -1 $O/r2_final_bench.json > profiles/r2_tf32_bench.json
for f in r2_tf32x3_bench r2_bf16_bench r2_reference_arm; do tail -1 $O/$f.json > profiles/$f.json; done
cp $O/kernel_pass_tf32.json profiles/r2_tf32_kernel_pass.json
python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
for f in profiles/r2f_ncu_*.txt; do echo "== $f"; grep -E "duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|derived" $f; done
