set -x
cd /root/repo
for i in 1 2; do
MUMPY_LANE_PRIO=0,0,0,0 timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
MUMPY_LANE_PRIO=-1,-1,0,-1 timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
MUMPY_LANE_PRIO=-2,-1,0,-1 timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
done
