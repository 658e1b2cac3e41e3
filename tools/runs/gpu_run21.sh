set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -x -k f32x2 > gpurun_out/r2_conv21.log 2>&1; echo "conv tests rc=$?"
tail -3 gpurun_out/r2_conv21.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes21_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench21_fp32.json 2> gpurun_out/r2_bench21_fp32.err; echo "rc=$?"
VAE2_FP32_TC_WGRAD_MIN_LANES=36 VAE2_BENCH_SHAPES=gpurun_out/r2_shapes21_fp32_w36.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench21_fp32_w36.json 2> gpurun_out/r2_bench21_fp32_w36.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench21*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
"
grep "36->36.*wgrad\|wgrad 36->36" gpurun_out/r2_shapes21_fp32_w36.txt | head -5 | cut -c1-170
grep "wgrad 270->270" gpurun_out/r2_shapes21_fp32.txt | head -3 | cut -c1-170
