mkdir -p gpurun_out
timeout 60 ./gpurun_scratch/ws_trace > gpurun_out/ws_trace2.log 2>&1; echo "trace rc=$?"
timeout 120 python tests/check_policy_tc.py > gpurun_out/r02_policy_tc_check5.log 2>&1; echo "check rc=$?"
WG_POLICY_TC=2 timeout 300 python -m pytest tests/test_cuda_policy.py -x -q -m gpu > gpurun_out/r02_pol_tests_ws.log 2>&1; echo "pytest(ws) rc=$?"; tail -3 gpurun_out/r02_pol_tests_ws.log
