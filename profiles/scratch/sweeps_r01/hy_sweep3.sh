mkdir -p gpurun_out
DLZ4_HY_SMEM_WARPS=6 DLZ4_HY_GL_WARPS=10 timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -x -q > gpurun_out/hy_pytest.log 2>&1
tail -3 gpurun_out/hy_pytest.log
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -x -q 2>&1 | tail -2
for cfg in "7 0" "7 1" "7 7" "6 10" "6 16" "6 22" "0 14"; do
  set -- $cfg
  echo "== smem_warps $1 gl_warps $2"
  DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed
done > gpurun_out/hy_sweep3.log 2>&1
cat gpurun_out/hy_sweep3.log
