set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "f32x3" > gpurun_out/r2_conv12.log 2>&1; echo "conv tests rc=$?"
tail -15 gpurun_out/r2_conv12.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes12_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench12_fp32.json 2> gpurun_out/r2_bench12_fp32.err; echo "rc=$?"
VAE2_FP32_TC_MIN_LANES=0 VAE2_BENCH_SHAPES=gpurun_out/r2_shapes12_fp32_all.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench12_fp32_all.json 2> gpurun_out/r2_bench12_fp32_all.err; echo "rc=$?"
VAE2_FP32_TC_MIN_LANES=0 VAE2_FP32_TC_WGRAD_MIN_LANES=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench12_fp32_allw.json 2> gpurun_out/r2_bench12_fp32_allw.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench12*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d['roofline']['kernel'], d['roofline']['share_of_step'])
    except Exception as e: print(f, 'ERR', e)
"
grep -h "f32x3" gpurun_out/r2_shapes12_fp32.txt | head -24
