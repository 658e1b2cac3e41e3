set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --no-bf16-path > gpurun_out/r2_bench31_n8_fp32.json 2> gpurun_out/r2_bench31_n8_fp32.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench31_n8_fp32.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench31_n8_fp32.json').read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['config'].get('syncbn'), d['config'].get('ddp'))
"
