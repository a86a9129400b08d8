mkdir -p gpurun_out
T=r02h
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/${T}_pytest_all.log | cut -c1-200
N="--clock-control none --profile-from-start off"
timeout 600 ncu --set full --import-source on $N -k regex:tc_wgrad -c 1 -o gpurun_out/${T}_wgrad16_config2 python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --import-source on $N -k regex:tc_chain_fwd -s 5 -c 3 -o gpurun_out/${T}_chains16_config2 python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_chains.log 2>&1; echo "ncu chains rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum $N --csv --log-file gpurun_out/${T}_launches_bwd2.csv python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_bwd2.log 2>&1; echo "launches bwd2 rc=$?"
