mkdir -p gpurun_out
N="--clock-control none --profile-from-start off"
timeout 300 python tools/profile_step.py --workload config3 --phase edge_bwd > gpurun_out/r02s_bwd3_default.log 2>&1; echo "bwd3 default rc=$?"; tail -1 gpurun_out/r02s_bwd3_default.log
CGNN_BWD_FUSED=1 timeout 300 python tools/profile_step.py --workload config3 --phase edge_bwd > gpurun_out/r02s_bwd3_fused.log 2>&1; echo "bwd3 fused rc=$?"; tail -1 gpurun_out/r02s_bwd3_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum $N --csv --log-file gpurun_out/r02s_launches_bwd3.csv python tools/profile_step.py --workload config3 --phase edge_bwd > gpurun_out/r02s_ncu_bwd3.log 2>&1; echo "launches bwd3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum $N --csv --log-file gpurun_out/r02s_launches_step3.csv python tools/profile_step.py --workload config3 > gpurun_out/r02s_ncu_step3.log 2>&1; echo "launches step3 rc=$?"
