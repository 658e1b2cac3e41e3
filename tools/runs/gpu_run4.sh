set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_gputest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest4.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest4.log
cp gpurun_out/parity_errors.jsonl gpurun_out/parity_errors_run4.jsonl
# the fp32-accurate tensor-core convolution (engine 2, split accumulators) inside the full nets
VAE2_FP32_TC=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_fp32" > gpurun_out/r2_gputest4_fp32tc.log 2>&1; echo "fp32tc rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/r2_gputest4_fp32tc.log
python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench4_bf16.json 2> gpurun_out/r2_bench4_bf16.err; echo "bench bf16 rc=$?"
VAE2_FP32_TC=1 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench4_fp32tc.json 2> gpurun_out/r2_bench4_fp32tc.err; echo "bench fp32tc rc=$?"
VAE2_D_STACK=1 timeout 900 python bench.py --steps 2 --warmup 1 --precision bf16 --workload w18_1024x2048 --no-cpu-baseline > gpurun_out/r2_bench4_1024.json 2> gpurun_out/r2_bench4_1024.err; echo "bench 1024x2048 rc=$?"
tail -3 gpurun_out/r2_bench4_1024.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench4*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['config']['per_gpu_batch'], d['hbm_peak_gb'], d.get('arena_gb'), d['gpu_launches'], d['roofline']['kernel'], d['roofline']['share_of_step'])
    except Exception as e: print(f, 'ERR', e)
"
