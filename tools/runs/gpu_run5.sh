set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_gputest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest5.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest5.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench5_default.json 2> gpurun_out/r2_bench5_default.err; echo "bench default rc=$?"
python bench.py --steps 2 --warmup 1 --precision bf16 --workload w48_473x473 --no-cpu-baseline > gpurun_out/r2_bench5_w48_473.json 2> gpurun_out/r2_bench5_w48_473.err; echo "bench w48 rc=$?"
tail -2 gpurun_out/r2_bench5_w48_473.err
NCU_ONLY="fwd,dgrad,wgrad" python tools/ncu_kernels.py fp32 > gpurun_out/ncu_plain_fp32.log 2>&1 && NCU_ONLY="fwd,dgrad,wgrad" ncu --set full --clock-control none --profile-from-start off -k regex:'conv_direct|wgrad_direct|conv_igemm|f32x3' -o gpurun_out/ncu_r2b_fp32 python tools/ncu_kernels.py fp32 > gpurun_out/ncu_run_fp32.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2b_fp32.ncu-rep --page raw --csv > gpurun_out/ncu_r2b_fp32_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/ncu_r2b_fp32.ncu-rep) -gt 30000000 ]; then rm gpurun_out/ncu_r2b_fp32.ncu-rep; fi
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench5*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],1), d['config']['per_gpu_batch'], d['hbm_peak_gb'], d.get('arena_gb'), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['share_of_step'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
"
du -sh gpurun_out
