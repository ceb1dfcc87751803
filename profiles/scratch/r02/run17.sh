set -x
cd /root/repo
(time timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/r02_bench_new1.txt 2>&1
tail -c 6000 gpurun_out/r02_bench_new1.txt
