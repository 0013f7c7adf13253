mkdir -p gpurun_out
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02_bench_n1e.json 2> gpurun_out/r02_bench_n1e.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
