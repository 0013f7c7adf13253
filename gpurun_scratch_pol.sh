mkdir -p gpurun_out
timeout 300 python profiles/r02_policy_tc_profile.py > gpurun_out/r02_plain_pol.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_act_tc -c 5 -f -o gpurun_out/r02_policy_tc_v4b python profiles/r02_policy_tc_profile.py > gpurun_out/r02_ncu_pol.log 2>&1
echo "ncu rc=$?"
