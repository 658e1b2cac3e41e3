set -x
mkdir -p gpurun_out
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes11_bf16.txt python bench.py --steps 2 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench11_bf16.json 2> gpurun_out/r2_bench11_bf16.err; echo "rc=$?"
VAE2_BENCH_SHAPES=gpurun_out/r2_shapes11_fp32.txt python bench.py --steps 2 --warmup 3 --no-bf16-path --no-cpu-baseline > gpurun_out/r2_bench11_fp32.json 2> gpurun_out/r2_bench11_fp32.err; echo "rc=$?"
head -40 gpurun_out/r2_shapes11_bf16.txt
