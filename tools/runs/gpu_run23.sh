set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x > gpurun_out/r2_conv23.log 2>&1; echo "conv tests rc=$?"
tail -3 gpurun_out/r2_conv23.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes23_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench23_fp32.json 2> gpurun_out/r2_bench23_fp32.err; echo "rc=$?"
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes23_bf16.txt timeout 600 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench23_bf16.json 2> gpurun_out/r2_bench23_bf16.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench23*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
"
grep "wgrad 256->18" gpurun_out/r2_shapes23_fp32.txt gpurun_out/r2_shapes23_bf16.txt | head -6 | cut -c1-200
