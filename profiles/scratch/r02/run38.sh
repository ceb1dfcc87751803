cd /root/repo
timeout 900 python -m pytest tests/test_gpu_frames.py -x -q -m gpu -k fuzz 2>&1 | tail -12 | cut -c1-250
