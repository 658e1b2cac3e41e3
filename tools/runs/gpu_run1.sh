set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2_gputest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest1.log
tail -5 gpurun_out/r2_gputest1.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.json
python tools/ncu_kernels.py bf16 > gpurun_out/ncu_plain_bf16.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc|wgrad_reduce|bn_fwd_fused|bn_bwd_fused|fuse_sum|fuse_bwd_up|elbo_terms' -o gpurun_out/ncu_r2a_bf16 python tools/ncu_kernels.py bf16 > gpurun_out/ncu_run_bf16.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_run_bf16.log
