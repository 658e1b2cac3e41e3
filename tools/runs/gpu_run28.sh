set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_gputest28.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_gputest28.log | head
timeout 900 python bench.py > gpurun_out/r2_bench28_default.json 2> gpurun_out/r2_bench28_default.err; echo "bench default rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench28_default.json').read().strip().splitlines()[-1])
print(round(d['value'],2), d['ms_per_step'], d['e2e'], d['gpu_launches'], d['hbm_peak_gb'], d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
b=d.get('bf16_path'); print(b and (round(b['value'],2), b['ms_per_step'], b['e2e']))
"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 13000 --launch-count 6400 --csv --log-file gpurun_out/r2_launches_fp32.csv python bench.py --steps 1 --warmup 1 --no-bf16-path --no-cpu-baseline > gpurun_out/ncu_launch28.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r2_launches_fp32.csv "ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 13000 --launch-count 6400 python bench.py --steps 1 --warmup 1 --no-bf16-path --no-cpu-baseline" > gpurun_out/r2_ncu_launch_summary_fp32.txt; head -20 gpurun_out/r2_ncu_launch_summary_fp32.txt
rm -f gpurun_out/r2_launches_fp32.csv
