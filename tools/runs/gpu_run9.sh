set -x
mkdir -p gpurun_out
VAE2_D_STACK=2 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline --batch 6 > gpurun_out/r2_bench9_fp32_b6s2.json 2> gpurun_out/r2_bench9_fp32_b6s2.err; echo "rc=$?"
VAE2_D_STACK=2 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline --batch 9 > gpurun_out/r2_bench9_bf16_b9s2.json 2> gpurun_out/r2_bench9_bf16_b9s2.err; echo "rc=$?"
python tools/bench_infer.py --precision bf16 --K 16 --clips 6 > gpurun_out/r2_infer9_bf16.json 2> gpurun_out/r2_infer9_bf16.err; echo "infer bf16 rc=$?"
tail -2 gpurun_out/r2_infer9_bf16.err
python tools/bench_infer.py --precision fp32 --K 16 --clips 4 --reference-draws 0 > gpurun_out/r2_infer9_fp32.json 2> gpurun_out/r2_infer9_fp32.err; echo "infer fp32 rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench9*.json')+glob.glob('gpurun_out/r2_infer9*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step', d.get('ms_per_clip_batch')), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d.get('arena_gb'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
