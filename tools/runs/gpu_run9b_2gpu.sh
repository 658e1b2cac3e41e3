set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity_full.py -m gpu -q -k "ddp or syncbn" > gpurun_out/r2_gputest9_dist.log 2>&1; echo "dist pytest rc=$?"
grep -E "PASS|FAIL|passed|failed" gpurun_out/r2_gputest9_dist.log | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench9_n2_fp32.json 2> gpurun_out/r2_bench9_n2_fp32.err; echo "bench n2 fp32 rc=$?"
tail -3 gpurun_out/r2_bench9_n2_fp32.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --precision bf16 > gpurun_out/r2_bench9_n2_bf16.json 2> gpurun_out/r2_bench9_n2_bf16.err; echo "bench n2 bf16 rc=$?"
tail -3 gpurun_out/r2_bench9_n2_bf16.err
timeout 60 tools/probes/umma_mn_shift_probe > gpurun_out/r2_probe_mn.txt 2>&1; echo "probe rc=$?"
grep -c exact gpurun_out/r2_probe_mn.txt; grep -c MISMATCH gpurun_out/r2_probe_mn.txt
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench9_n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('per_gpu_batch'), d.get('hbm_peak_gb'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
"
