set -x
cd /root/repo
timeout 600 python bench.py --no-kernels --no-fp16 --steps 20 2>&1 | tail -1 > gpurun_out/scale_n1.json
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-kernels --no-fp16 2>&1 | tail -1 > gpurun_out/scale_n$n.json
done
for n in 1 2 4 8; do cut -c1-160 gpurun_out/scale_n$n.json; done
