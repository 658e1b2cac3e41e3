set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "p2p" > gpurun_out/r2_gputest17_dist.log 2>&1; echo "dist pytest rc=$?"
grep -E "PASS|FAIL|passed|failed|peer-memory|Error|error" gpurun_out/r2_gputest17_dist.log | head -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench17_n2_fp32_p2p.json 2> gpurun_out/r2_bench17_n2_fp32_p2p.err; echo "bench n2 fp32 p2p rc=$?"
tail -3 gpurun_out/r2_bench17_n2_fp32_p2p.err
VAE2_SYNCBN_P2P=0 timeout 600 $TR --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench17_n2_fp32_nccl.json 2> gpurun_out/r2_bench17_n2_fp32_nccl.err; echo "bench n2 fp32 nccl rc=$?"
timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --precision bf16 > gpurun_out/r2_bench17_n2_bf16_p2p.json 2> gpurun_out/r2_bench17_n2_bf16_p2p.err; echo "bench n2 bf16 p2p rc=$?"
tail -3 gpurun_out/r2_bench17_n2_bf16_p2p.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench17_n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('per_gpu_batch'), d['config'].get('syncbn'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
"
