cd /root/repo
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand 2>&1 | cut -c1-130
timeout 300 python divortio-lz4_b200/tools/config_bench.py 1024 262144 2>&1 | head -4
