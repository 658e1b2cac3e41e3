set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x > gpurun_out/r2_conv26.log 2>&1; echo "conv tests rc=$?"
tail -12 gpurun_out/r2_conv26.log
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes26_bf16.txt timeout 600 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench26_bf16.json 2> gpurun_out/r2_bench26_bf16.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench26_bf16.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench26_fp32.json 2> gpurun_out/r2_bench26_fp32.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench26*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), d.get('ms_per_step'))
        for r in d['kernel_shares'][:12]: print('  %-66s share %.3f ms %.1f n %d hbm %.3f tensor %.3f'%(r['kernel'],r['share'],r['ms'],r['n'],r['hbm_frac'],r['tensor_frac']))
    except Exception as e: print(f, 'ERR', e)
"
grep "270->3" gpurun_out/r2_shapes26_bf16.txt | cut -c1-170
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -q -x -k "bf16" > gpurun_out/r2_parity26.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r2_parity26.log
