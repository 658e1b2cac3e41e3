set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x > gpurun_out/r2_conv20.log 2>&1; echo "conv tests rc=$?"
tail -5 gpurun_out/r2_conv20.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes20_fp32.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench20_fp32.json 2> gpurun_out/r2_bench20_fp32.err; echo "rc=$?"
VAE2_FP32_TC_WGRAD_NARROW=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench20_fp32_nonarrow.json 2> gpurun_out/r2_bench20_fp32_nonarrow.err; echo "rc=$?"
VAE2_WGRAD_DUAL=0 VAE2_FP32_TC_WGRAD_NARROW=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench20_fp32_nodual.json 2> gpurun_out/r2_bench20_fp32_nodual.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench20*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'), d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d['roofline']['kernel'], d['roofline']['share_of_step'])
        for r in d['kernel_shares'][:10]: print('  %-66s share %.3f ms %.1f n %d hbm %.3f tensor %.3f'%(r['kernel'],r['share'],r['ms'],r['n'],r['hbm_frac'],r['tensor_frac']))
    except Exception as e: print(f, 'ERR', e)
"
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -q -x -k "fp32 or golden or w18 or W18" > gpurun_out/r2_parity20.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r2_parity20.log
