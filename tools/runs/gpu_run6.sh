set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_gputest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest6.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest6.log
python tools/microbench_conv.py bf16 10 > gpurun_out/r2_micro6_bf16_ctas2.txt 2>&1
VAE2_WGRAD_CTAS=1 python tools/microbench_conv.py bf16 10 > gpurun_out/r2_micro6_bf16_ctas1.txt 2>&1
paste -d'\n' gpurun_out/r2_micro6_bf16_ctas2.txt gpurun_out/r2_micro6_bf16_ctas1.txt | cut -c1-220
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6_default.json 2> gpurun_out/r2_bench6_default.err; echo "bench default rc=$?"
tail -2 gpurun_out/r2_bench6_default.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench6*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],1), d['config']['per_gpu_batch'], d['hbm_peak_gb'], d.get('arena_gb'), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['share_of_step'],3), 'e2e', round(d['e2e']['value'],1), 'bf16', d.get('bf16_path',{}).get('value'))
        for r in d['kernel_shares'][:8]: print('   ', r['kernel'], round(100*r['share'],1), round(r['ms'],1), r['n'])
    except Exception as e: print(f, 'ERR', e)
"
