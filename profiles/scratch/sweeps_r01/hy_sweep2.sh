mkdir -p gpurun_out
for cfg in "0 1" "0 7" "0 14" "0 28" "1 0"; do
  set -- $cfg
  echo "== smem_warps $1 gl_warps $2"
  DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 300 python divortio-lz4_b200/tools/kbench.py 128 65536 log
done > gpurun_out/hy_sweep2.log 2>&1
cat gpurun_out/hy_sweep2.log
DLZ4_HY_SMEM_WARPS=0 DLZ4_HY_GL_WARPS=7 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16h -c 1 -o gpurun_out/prof_hy_gl7 -f python divortio-lz4_b200/tools/prof_one.py log 64 > gpurun_out/ncu_hy.log 2>&1
tail -3 gpurun_out/ncu_hy.log
