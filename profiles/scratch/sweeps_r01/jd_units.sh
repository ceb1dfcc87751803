for u in 8 16 32; do echo "== unit MiB $u"; DLZ4_JD_UNIT_MIB=$u timeout 600 python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --no-cpu 2>&1 | grep "cc=0" | cut -c1-48,180-300; done
timeout 300 python divortio-lz4_b200/tools/frame_bench.py log 64 --no-cpu 2>&1 | grep "cc=0" | cut -c1-48,180-300
timeout 900 python -m pytest tests/test_gpu_jump_decode.py tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -3
