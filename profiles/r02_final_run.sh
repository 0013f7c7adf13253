# final single-GPU evidence of round 2 (one gpurun call; every profiled command first exits 0 without ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02_final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference arm rc=$?"
H="python bench.py --steps 20 --warmup 5 --no-sub --no-e2e --no-cpu-baseline"
$H > gpurun_out/r02_plain_headline.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv $H > gpurun_out/r02_ncu_ll.log 2>&1
echo "launch list rc=$?"
C5="python bench.py --config 5 --steps 64 --warmup 32 --no-e2e --no-cpu-baseline"
$C5 > gpurun_out/r02_plain_cfg5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg5.csv $C5 > gpurun_out/r02_ncu_ll5.log 2>&1
echo "cfg5 launch list rc=$?"
P="python profiles/r02_policy_tc_profile.py 2"
$P > gpurun_out/r02_plain_pol.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:policy_act_ws -c 4 -f -o gpurun_out/r02_policy_ws $P > gpurun_out/r02_ncu_pol.log 2>&1
echo "policy full rc=$?"
