# ncu evidence for round 2 (one gpurun call; every profiled command first exits 0 without ncu)
mkdir -p gpurun_out
H="python bench.py --steps 20 --warmup 5 --no-sub --no-e2e --no-cpu-baseline"
C4="python bench.py --config 4 --body quad_chain --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
U4="python bench.py --config 4 --body quad --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
$H > gpurun_out/r02_plain_headline.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv $H > gpurun_out/r02_ncu_ll.log 2>&1
echo "launch list rc=$?"
$H > gpurun_out/r02_plain_headline2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_static_packed -s 270 -c 2 -f -o gpurun_out/r02_step_balance_steady $H > gpurun_out/r02_ncu_full.log 2>&1
echo "headline full rc=$?"
$C4 > gpurun_out/r02_plain_chain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_units -s 250 -c 1 -f -o gpurun_out/r02_step_units_linked $C4 > gpurun_out/r02_ncu_chain.log 2>&1
echo "chain full rc=$?"
$U4 > gpurun_out/r02_plain_units.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_units -s 250 -c 1 -f -o gpurun_out/r02_step_units $U4 > gpurun_out/r02_ncu_units.log 2>&1
echo "units full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
# (the units capture under profiles/ was retaken after the packed-pair arithmetic: gpurun_scratch/run_o.sh = the U4 block above)
