set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "f32x3" > gpurun_out/r2_conv13.log 2>&1; echo "conv tests rc=$?"
tail -5 gpurun_out/r2_conv13.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes13_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench13_fp32.json 2> gpurun_out/r2_bench13_fp32.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench13*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d['roofline']['kernel'], d['roofline']['share_of_step'])
        for r in d['kernel_shares'][:10]: print('  %-60s share %.3f ms %.1f n %d hbm %.3f tensor %.3f'%(r['kernel'],r['share'],r['ms'],r['n'],r['hbm_frac'],r['tensor_frac']))
    except Exception as e: print(f, 'ERR', e)
"
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_gputest13.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_gputest13.log | head -20
