#!/bin/bash
# ncu captures for profiles/: (A) full set + source for the fused set-abstraction kernels of one step, (B) speed-of-light / memory /
# occupancy sections for one step's worth of the other kernel families.  usage: tools/gpu_ncu.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"sa_fwd_kernel|sa_bwd_kernel|sa_bwd_pre_kernel|sa_u_kernel|sa_u_bwd_kernel|sa_final_kernel" \
    --launch-skip 76 --launch-count 38 -f -o $OUT/${TAG}_sa $CMD > $OUT/${TAG}_ncuA.log 2>&1
tail -2 $OUT/${TAG}_ncuA.log
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --clock-control none \
    -k regex:"wgrad_tc_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel|bn_act_split_kernel|conv_gemm_tc_kernel|fps_chain_kernel|heads_linear|interp_|three_nn|upsample2x|im2col_split|fin_finalize|prep_weight_batch|rows_mean|bn_relu_max|sa_gather|sa_scatter|adam_flat" \
    --launch-skip 1300 --launch-count 650 -f -o $OUT/${TAG}_fam $CMD > $OUT/${TAG}_ncuB.log 2>&1
tail -2 $OUT/${TAG}_ncuB.log
ls -la $OUT/${TAG}_sa.ncu-rep $OUT/${TAG}_fam.ncu-rep
