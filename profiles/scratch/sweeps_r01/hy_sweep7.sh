mkdir -p gpurun_out
DLZ4_HY_SMEM_WARPS=4 DLZ4_HY_GL_WARPS=24 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -x -q 2>&1 | tail -2
for cfg in "7 0" "4 24" "4 28" "6 12"; do
  set -- $cfg
  echo "== smem_warps $1 gl_warps $2"
  DLZ4_HY_PERSIST=0 DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed
done > gpurun_out/hy_sweep7.log 2>&1
cat gpurun_out/hy_sweep7.log
