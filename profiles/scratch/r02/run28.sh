set -x
cd /root/repo
nvidia-smi -L | wc -l; nproc; free -g | head -2
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3) > gpurun_out/r02_bench_n8.txt 2>&1
grep -v Warning gpurun_out/r02_bench_n8.txt | tail -c 7000
