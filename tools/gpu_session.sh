# scratch: the command list of the last gpurun session (development aid)
cd /root/repo
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final.json
