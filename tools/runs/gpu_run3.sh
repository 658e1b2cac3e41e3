set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2_gputest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest3.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest3.log
python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench3_bf16.json 2> gpurun_out/r2_bench3_bf16.err; echo "bench bf16 rc=$?"
tail -3 gpurun_out/r2_bench3_bf16.err
VAE2_D_STACK=2 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline --batch 8 > gpurun_out/r2_bench3_bf16_b8s2.json 2> gpurun_out/r2_bench3_bf16_b8s2.err; echo "bench bf16 b8 stack2 rc=$?"
python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench3_fp32.json 2> gpurun_out/r2_bench3_fp32.err; echo "bench fp32 rc=$?"
tail -3 gpurun_out/r2_bench3_fp32.err
VAE2_D_STACK=2 timeout 900 python bench.py --steps 2 --warmup 1 --precision bf16 --workload w18_1024x2048 --no-cpu-baseline > gpurun_out/r2_bench3_1024.json 2> gpurun_out/r2_bench3_1024.err; echo "bench 1024x2048 rc=$?"
tail -5 gpurun_out/r2_bench3_1024.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench3*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['config']['per_gpu_batch'], d['hbm_peak_gb'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['share_of_step'])
    except Exception as e: print(f, 'ERR', e)
"
du -sh gpurun_out
