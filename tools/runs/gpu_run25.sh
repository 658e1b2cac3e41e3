set -x
mkdir -p gpurun_out
export NCU_ONLY="wgrad 18,wgrad 64->64,wgrad 256,wgrad 36,fwd 18->18,fwd 64->64"
timeout 300 python tools/ncu_kernels.py bf16 > gpurun_out/ncu_plain25.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'wgrad_halo|wgrad_tc_kernel|wgrad_reduce|conv_tc' -o gpurun_out/ncu_r2f_bf16 python tools/ncu_kernels.py bf16 > gpurun_out/ncu_run25.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2f_bf16.ncu-rep --page raw --csv > gpurun_out/ncu_r2f_bf16_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/ncu_r2f_bf16.ncu-rep) -gt 30000000 ]; then rm gpurun_out/ncu_r2f_bf16.ncu-rep; fi
unset NCU_ONLY
timeout 600 python bench.py --steps 3 --warmup 3 --precision bf16 --workload w48_473x473 --no-cpu-baseline > gpurun_out/r2_bench25_w48_473.json 2> gpurun_out/r2_bench25_w48_473.err; echo "w48 rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 --precision bf16 --workload w18_1024x2048 --no-cpu-baseline > gpurun_out/r2_bench25_1024.json 2> gpurun_out/r2_bench25_1024.err; echo "1024 rc=$?"
timeout 600 python tools/bench_infer.py --precision bf16 --K 16 --clips 6 > gpurun_out/r2_infer25_bf16.json 2> gpurun_out/r2_infer25_bf16.err; echo "infer bf16 rc=$?"
timeout 600 python tools/bench_infer.py --precision fp32 --K 16 --clips 4 --reference-draws 0 > gpurun_out/r2_infer25_fp32.json 2> gpurun_out/r2_infer25_fp32.err; echo "infer fp32 rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench25*.json')+glob.glob('gpurun_out/r2_infer25*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step', d.get('ms_per_clip_batch')), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'))
    except Exception as e: print(f, 'ERR', e)
"
du -sh gpurun_out
