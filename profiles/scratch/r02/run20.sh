set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "page_locked or multi_frame or config3" > gpurun_out/r02_async_tests.txt 2>&1
tail -15 gpurun_out/r02_async_tests.txt
(time timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline) > gpurun_out/r02_bench_new3.txt 2>&1
tail -c 4000 gpurun_out/r02_bench_new3.txt
