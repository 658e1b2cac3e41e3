set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench30_default.json 2> gpurun_out/r2_bench30_default.err; echo "bench default rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench30_default.json').read().strip().splitlines()[-1])
print(round(d['value'],2), d['ms_per_step'], d['e2e'], d['gpu_launches'], d['hbm_peak_gb'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('traffic_source'), d['roofline'].get('tensor_issued'), d.get('cpu_baseline',{}).get('value'), d['clocks'])
b=d.get('bf16_path'); print(b and (round(b['value'],2), b['ms_per_step'], b['e2e'], b['roofline']['kernel'], b['roofline']['frac']))
"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke30.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke30.log
