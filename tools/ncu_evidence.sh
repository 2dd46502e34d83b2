#!/bin/bash
# `ncu --set full` captures of the final kernels (one launch each, after warm-up) -> gpurun_out/r2_ncu_full_*.txt
# (run on the GPU box: gpurun -- 'bash tools/ncu_evidence.sh'; copy the summaries into profiles/)
set -x
NCU="ncu --set full --clock-control none --import-source on"
for c in last800_320 pyr160_ln_gelu head32_32_elu; do
  timeout 300 $NCU -k regex:gwd_tapgemm -s 5 -c 1 python tools/bench_gemm.py $c > gpurun_out/r2_ncu_full_tapgemm_$c.txt 2>&1
done
for c in enc_b16_l300 enc_b64_l1200; do
  timeout 300 $NCU -k regex:gwd_attention -s 5 -c 1 python tools/bench_attention.py $c > gpurun_out/r2_ncu_full_attention_$c.txt 2>&1
done
# the training step's heaviest non-GEMM kernels, one launch each
timeout 600 $NCU -k regex:'gwd_wgrad_tc_kernel|gwd_layernorm_bwd_kernel|gwd_ref_bwd_kernel|gwd_diffuse_wgrad_kernel|gwd_col2im3x3_s2_kernel|gwd_im2col3x3_s2_kernel' \
  --nvtx --nvtx-include "train_step/" -c 12 python tools/profile_train_step.py > gpurun_out/r2_ncu_full_train_kernels.txt 2>&1
ls -la gpurun_out/r2_ncu_full_*
