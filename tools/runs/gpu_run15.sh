set -x
mkdir -p gpurun_out
export NCU_ONLY="f32x3,f32x2"
timeout 600 python tools/ncu_kernels.py fp32 > gpurun_out/ncu_plain15.log 2>&1; echo "plain rc=$?"
tail -3 gpurun_out/ncu_plain15.log
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'f32x3|wgrad_halo|wgrad_tc_kernel|split_planes|wgrad_reduce' -o gpurun_out/ncu_r2d_fp32 python tools/ncu_kernels.py fp32 > gpurun_out/ncu_run15.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2d_fp32.ncu-rep --page raw --csv > gpurun_out/ncu_r2d_fp32_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_r2d_fp32.ncu-rep --page details --csv > gpurun_out/ncu_r2d_fp32_details.csv 2>/dev/null
ls -la gpurun_out/ncu_r2d_fp32*
if [ $(stat -c %s gpurun_out/ncu_r2d_fp32.ncu-rep) -gt 40000000 ]; then rm gpurun_out/ncu_r2d_fp32.ncu-rep; fi
du -sh gpurun_out
