cd /root/repo
for w in heads cxc32; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_gemm_$w python tools/one_gemm.py $w mixed > /dev/null 2>&1
ncu -i gpurun_out/r2_gemm_$w.ncu-rep --page raw --csv > gpurun_out/r2_gemm_${w}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_gemm_$w.ncu-rep --page source --csv > gpurun_out/r2_gemm_${w}_src.csv 2>/dev/null
done
