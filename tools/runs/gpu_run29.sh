set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-bf16-path --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r2_bench29_plain.json 2> gpurun_out/r2_bench29_plain.err; echo "plain rc=$?"
VAE2_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_fp32.csv $CMD > gpurun_out/ncu_launch29.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r2_launches_fp32.csv "VAE2_BENCH_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off $CMD   (the ONE timed step, CUDA graphs on)" > gpurun_out/r2_ncu_launch_summary_fp32.txt; head -32 gpurun_out/r2_ncu_launch_summary_fp32.txt
ls -la gpurun_out/r2_launches_fp32.csv; rm -f gpurun_out/r2_launches_fp32.csv
