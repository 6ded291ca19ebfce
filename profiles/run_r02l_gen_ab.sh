python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv_umma" 2>&1 | tail -3
for L in ds2.c0 ds2.c1 bg.c0 bg.c1 us0.c0 us0.c1; do
  a=$(DRAM_CONV_V4=0 python tests/micro_conv.py 8 $L 2>&1 | tail -1); b=$(python tests/micro_conv.py 8 $L 2>&1 | tail -1)
  echo "$L single-SM: $a"; echo "$L SM pairs : $b"
done
