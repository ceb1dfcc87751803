mkdir -p gpurun_out
export DLZ4_HY_SMEM_WARPS=4 DLZ4_HY_GL_WARPS=24 DLZ4_HY_PERSIST=0
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16h -c 1 -o gpurun_out/prof_hy_4_24_log -f python divortio-lz4_b200/tools/prof_one.py log 512 > gpurun_out/ncu_hy2.log 2>&1
tail -2 gpurun_out/ncu_hy2.log
