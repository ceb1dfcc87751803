mkdir -p gpurun_out
for w in 512 384 256; do for s in 0 64 256; do
  echo "== warm KiB $w seg KiB $s (0 = auto)"
  if [ $s = 0 ]; then unset DLZ4_SEG_KIB; else export DLZ4_SEG_KIB=$s; fi
  DLZ4_SEG_WARM_KIB=$w timeout 300 python divortio-lz4_b200/tools/frame_bench.py log 64 --no-cpu 2>&1 | grep "4194304 linked      cc=0" | cut -c1-175
  DLZ4_SEG_WARM_KIB=$w timeout 300 python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --no-cpu 2>&1 | grep "4194304 linked      cc=0" | cut -c1-175
done; done > gpurun_out/seg_sweep.log 2>&1
cat gpurun_out/seg_sweep.log
