set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x > gpurun_out/r2_conv10.log 2>&1; echo "conv tests rc=$?"
tail -15 gpurun_out/r2_conv10.log
VAE2_WGRAD_HALO=0 timeout 300 python tools/microbench_conv.py bf16 > gpurun_out/r2_mb10_nohalo.txt 2>&1
VAE2_WGRAD_HALO=1 timeout 300 python tools/microbench_conv.py bf16 > gpurun_out/r2_mb10_halo.txt 2>&1
paste -d'\n' gpurun_out/r2_mb10_nohalo.txt gpurun_out/r2_mb10_halo.txt | grep -o "^.\{60\}\|wgrad.*" | paste - - | head -60
python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench10_bf16.json 2> gpurun_out/r2_bench10_bf16.err; echo "rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench10_fp32.json 2> gpurun_out/r2_bench10_fp32.err; echo "rc=$?"
VAE2_WGRAD_HALO=0 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench10_bf16_nohalo.json 2> gpurun_out/r2_bench10_bf16_nohalo.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench10*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d['roofline']['kernel'], d['roofline']['share_of_step'])
    except Exception as e: print(f, 'ERR', e)
"
bash tools/runs/gpu_run9.sh
