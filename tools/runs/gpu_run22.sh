set -x
mkdir -p gpurun_out
export NCU_ONLY="f32x2"
timeout 600 python tools/ncu_kernels.py fp32 > gpurun_out/ncu_plain22.log 2>&1; echo "plain rc=$?"
tail -3 gpurun_out/ncu_plain22.log
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:'wgrad_halo|wgrad_tc_kernel|split_planes|wgrad_reduce' -o gpurun_out/ncu_r2e_fp32_wgrad python tools/ncu_kernels.py fp32 > gpurun_out/ncu_run22.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2e_fp32_wgrad.ncu-rep --page raw --csv > gpurun_out/ncu_r2e_fp32_wgrad_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_r2e*
if [ $(stat -c %s gpurun_out/ncu_r2e_fp32_wgrad.ncu-rep) -gt 40000000 ]; then rm gpurun_out/ncu_r2e_fp32_wgrad.ncu-rep; fi
unset NCU_ONLY
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_gputest22.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_gputest22.log | head -20
timeout 900 python bench.py > gpurun_out/r2_bench22_default.json 2> gpurun_out/r2_bench22_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench22_ref.json 2> gpurun_out/r2_bench22_ref.err; echo "bench ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench22_default.json').read().strip().splitlines()[-1])
print(round(d['value'],2), d['ms_per_step'], d['e2e'], d['gpu_launches'], d['hbm_peak_gb'], d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
b=d.get('bf16_path'); print(b and (round(b['value'],2), b['ms_per_step'], b['e2e']))
print(open('gpurun_out/r2_bench22_ref.json').read()[:600])
"
du -sh gpurun_out
