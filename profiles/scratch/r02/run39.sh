cd /root/repo
(time timeout 900 python bench.py --steps 5 --warmup 3 --no-configs) > gpurun_out/r02_bench_check.txt 2>&1
tail -c 2500 gpurun_out/r02_bench_check.txt
