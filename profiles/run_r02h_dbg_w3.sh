for L in us2.c0 us2.c1; do for D in 0 1 2 3; do echo "== $L DBG=$D"; DRAM_CONV_DBG=$D python tests/micro_conv.py 8 $L 2>&1 | tail -1; done; done > gpurun_out/r02h_dbg_w3.txt 2>&1
cat gpurun_out/r02h_dbg_w3.txt
