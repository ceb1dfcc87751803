cd /root/repo
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-220
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python divortio-lz4_b200/tools/prof_one.py log 8 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -40 | cut -c1-260
