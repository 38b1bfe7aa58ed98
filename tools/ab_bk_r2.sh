for v in product bk_ls640 bk_ls512 bk_ls320x2 bk_ls256x2 bk_f64x10 bk_ls640s4; do
  if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
  timeout 200 python tests/perf/time_bickley.py 2>&1 | grep -v Warning
done > gpurun_out/r2g_ab_bickley.txt 2>&1
cat gpurun_out/r2g_ab_bickley.txt
