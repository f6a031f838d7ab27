set -x
cd /root/repo
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final.json
timeout 600 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown.txt 2>&1; head -3 gpurun_out/op_breakdown.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
