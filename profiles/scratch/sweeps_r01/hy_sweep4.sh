mkdir -p gpurun_out
for cfg in "0 18" "0 22" "0 28" "2 18" "3 16" "4 14" "4 18" "5 12" "6 12" "6 8"; do
  set -- $cfg
  echo "== smem_warps $1 gl_warps $2"
  DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed
done > gpurun_out/hy_sweep4.log 2>&1
cat gpurun_out/hy_sweep4.log
