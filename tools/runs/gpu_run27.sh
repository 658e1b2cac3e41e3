set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_tc.py -q -x > gpurun_out/r2_kern27.log 2>&1; echo "kernel tests rc=$?"
tail -4 gpurun_out/r2_kern27.log
timeout 600 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench27_bf16.json 2> gpurun_out/r2_bench27_bf16.err; echo "rc=$?"
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes27_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench27_fp32.json 2> gpurun_out/r2_bench27_fp32.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench27*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
"
grep "bias_grad" gpurun_out/r2_shapes27_fp32.txt | cut -c1-170
