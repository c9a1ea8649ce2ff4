#!/bin/bash
# ncu captures for profiles/ (one profiled eager step between cudaProfilerStart/Stop, tools/profile_step.py):
#  (A) --set full for the fused set-abstraction kernels of one extractor,
#  (B) speed-of-light + memory sections for the first launches of the other kernel families.
# The .ncu-rep files are turned into text tables ON THE BOX (tools/ncu_table.py) and large reports are deleted: gpurun_out/ may
# carry at most 64 MiB back.  usage: tools/gpu_ncu.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/profile_step.py --warmup 2 --steps 1"
timeout 600 ncu --set full --clock-control none --profile-from-start off \
    -k regex:"sa_fwd_kernel|sa_bwd_kernel|sa_bwd_pre_kernel|sa_u_kernel|sa_u_bwd_kernel|sa_final_kernel" \
    --launch-count 19 -f -o $OUT/${TAG}_sa $CMD > $OUT/${TAG}_ncuA.log 2>&1
tail -2 $OUT/${TAG}_ncuA.log
python tools/ncu_table.py $OUT/${TAG}_sa.ncu-rep > $OUT/${TAG}_ncu_sa_fused.txt 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --profile-from-start off \
    -k regex:"wgrad_tc_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel|bn_act_split_kernel|conv_gemm_tc_kernel|fps_chain_kernel|heads_linear|interp_|three_nn|upsample2x|im2col_split|fin_finalize|prep_weight_batch|rows_mean|bn_relu_max|sa_gather|sa_scatter" \
    --launch-count 260 -f -o $OUT/${TAG}_fam $CMD > $OUT/${TAG}_ncuB.log 2>&1
tail -2 $OUT/${TAG}_ncuB.log
python tools/ncu_table.py $OUT/${TAG}_fam.ncu-rep --group > $OUT/${TAG}_ncu_families.txt 2>&1
python tools/ncu_table.py $OUT/${TAG}_fam.ncu-rep > $OUT/${TAG}_ncu_families_all.txt 2>&1
ls -la $OUT/${TAG}_sa.ncu-rep $OUT/${TAG}_fam.ncu-rep
for f in $OUT/${TAG}_sa.ncu-rep $OUT/${TAG}_fam.ncu-rep; do
  sz=$(stat -c %s $f 2>/dev/null || echo 0)
  if [ "$sz" -gt 20000000 ]; then rm -f $f; echo "removed $f ($sz bytes)"; fi
done
du -sh $OUT
