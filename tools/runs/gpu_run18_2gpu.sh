set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
VAE2_BENCH_DDP_BCAST=0 timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench18_n2_fp32_nobcast.json 2> gpurun_out/r2_bench18_a.err; echo "rc=$?"
VAE2_BENCH_DDP_BCAST=0 VAE2_BENCH_DDP_BUCKET_VIEW=1 timeout 600 $TR --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench18_n2_fp32_nobcast_view.json 2> gpurun_out/r2_bench18_b.err; echo "rc=$?"
VAE2_BENCH_DDP_BCAST=0 VAE2_BENCH_DDP_BUCKET_VIEW=1 timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path --no-graphs > gpurun_out/r2_bench18_n2_fp32_nobcast_view_eager.json 2> gpurun_out/r2_bench18_c.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench18_a.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench18_n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('per_gpu_batch'), d['config'].get('syncbn'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
"
