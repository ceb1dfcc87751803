mkdir -p gpurun_out
for cfg in "0 32" "4 24" "4 28" "2 26" "3 22" "4 20"; do
  set -- $cfg
  echo "== smem_warps $1 gl_warps $2"
  DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed
done > gpurun_out/hy_sweep5.log 2>&1
cat gpurun_out/hy_sweep5.log
for cfg in "0 28" "4 20"; do
  set -- $cfg
  echo "== ncu smem_warps $1 gl_warps $2"
  DLZ4_HY_SMEM_WARPS=$1 DLZ4_HY_GL_WARPS=$2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_compress_fresh16h -c 1 python divortio-lz4_b200/tools/prof_one.py mixed 1024 2>&1 | grep -E "dram__|lts__|l1tex__|gpu__time|smsp__"
done > gpurun_out/hy_ncu5.log 2>&1
cat gpurun_out/hy_ncu5.log
