set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cap() {  # $1 = one_kernel.py case, $2 = kernel base-name regex
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 --launch-skip 2 -f -o gpurun_out/r1_k_$1 python tools/one_kernel.py $1 > gpurun_out/prof_k_$1.log 2>&1
  ncu -i gpurun_out/r1_k_$1.ncu-rep --page raw --csv > gpurun_out/r1_k_$1_raw.csv 2>/dev/null
}
cap attn attn_fused_kernel
cap ln layernorm_kernel
cap ln_bwd layernorm_bwd_kernel
cap resid_bwd resid_branch_bwd_kernel
cap softmax softmax_rows
cap to_planes to_planes_vec_kernel
ls -la gpurun_out | grep "r1_k_.*raw" | head -20
