set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2_gputest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest2.log
tail -30 gpurun_out/r2_gputest2.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
python tools/ncu_kernels.py bf16 > gpurun_out/ncu_plain_bf16.log 2>&1 && ncu --set full --clock-control none --profile-from-start off -k regex:'conv_tc|wgrad_tc|wgrad_reduce|bn_fwd_fused|bn_bwd_fused|fuse_sum|fuse_bwd_up|elbo_terms' -o gpurun_out/ncu_r2a_bf16 python tools/ncu_kernels.py bf16 > gpurun_out/ncu_run_bf16.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2a_bf16.ncu-rep --page raw --csv > gpurun_out/ncu_r2a_bf16_raw.csv 2>/dev/null
ls -la gpurun_out
# keep the report only if it fits the 64 MiB return limit with room to spare
if [ $(stat -c %s gpurun_out/ncu_r2a_bf16.ncu-rep) -gt 40000000 ]; then rm gpurun_out/ncu_r2a_bf16.ncu-rep; fi
du -sh gpurun_out
