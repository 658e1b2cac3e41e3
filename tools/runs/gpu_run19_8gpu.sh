set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench19_n8_fp32_p2p.json 2> gpurun_out/r2_bench19_n8_fp32.err; echo "rc=$?"
tail -4 gpurun_out/r2_bench19_n8_fp32.err
timeout 420 $TR --master-port 29534 bench.py --gpus 8 --steps 3 --warmup 3 --precision bf16 > gpurun_out/r2_bench19_n8_bf16_p2p.json 2> gpurun_out/r2_bench19_n8_bf16.err; echo "rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench19_n8*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('per_gpu_batch'), d['config'].get('syncbn'), d.get('gpu_launches'))
        for r in d['kernel_shares'][:8]: print('  %-60s share %.3f ms %.1f n %d'%(r['kernel'],r['share'],r['ms'],r['n']))
    except Exception as e: print(f, 'ERR', e)
"
