set -x
cd /root/repo
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "two_real" > gpurun_out/r02_two_gpu_test.txt 2>&1
tail -3 gpurun_out/r02_two_gpu_test.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r02_bench_n2.txt 2>&1
tail -c 5000 gpurun_out/r02_bench_n2.txt
