set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2_gputest24_dist.log 2>&1; echo "dist pytest rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_gputest24_dist.log | head -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-bf16-path > gpurun_out/r2_bench24_n2_fp32.json 2> gpurun_out/r2_bench24_n2_fp32.err; echo "bench n2 fp32 rc=$?"
timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 --precision bf16 > gpurun_out/r2_bench24_n2_bf16.json 2> gpurun_out/r2_bench24_n2_bf16.err; echo "bench n2 bf16 rc=$?"
timeout 600 $TR --master-port 29536 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench24_n2_ref.json 2> gpurun_out/r2_bench24_n2_ref.err; echo "bench n2 ref rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench24_n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('per_gpu_batch'), d['config'].get('syncbn'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
"
