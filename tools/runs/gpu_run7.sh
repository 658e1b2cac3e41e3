set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_gputest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest7.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest7.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench7_default.json 2> gpurun_out/r2_bench7_default.err; echo "bench default rc=$?"
tail -2 gpurun_out/r2_bench7_default.err
NCU_ONLY="fuse,elbo,bn " python tools/ncu_kernels.py bf16 > gpurun_out/ncu_plain7.log 2>&1 && NCU_ONLY="fuse,elbo,bn " ncu --set full --clock-control none --profile-from-start off -k regex:'bn_fwd_fused|bn_bwd_fused|fuse_sum|fuse_bwd_up|elbo_terms' -o gpurun_out/ncu_r2c_bf16 python tools/ncu_kernels.py bf16 > gpurun_out/ncu_run7.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2c_bf16.ncu-rep --page raw --csv > gpurun_out/ncu_r2c_bf16_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/ncu_r2c_bf16.ncu-rep) -gt 30000000 ]; then rm gpurun_out/ncu_r2c_bf16.ncu-rep; fi
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench7*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],1), d['config']['per_gpu_batch'], d['hbm_peak_gb'], d.get('arena_gb'), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['share_of_step'],3), 'e2e', round(d['e2e']['value'],1), 'bf16', d.get('bf16_path',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))
        for r in d['kernel_shares'][:10]: print('   ', r['kernel'], round(100*r['share'],1), round(r['ms'],1), r['n'], 'hbm', round(100*r['hbm_frac'],1), 'tensor', round(100*r['tensor_frac'],2))
    except Exception as e: print(f, 'ERR', e)
"
