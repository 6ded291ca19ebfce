export DRAM_CONV_V4=1
for L in us2.c1 us2.c0; do for D in 0 4 7; do echo "== $L DBG=$D"; DRAM_CONV_PROF=1 DRAM_CONV_DBG=$D python tests/micro_conv.py 8 $L 2>&1 | grep -E "total|fwd4 prof" | tail -4; done; done > gpurun_out/r02c_dbg_pairs2.txt 2>&1
cat gpurun_out/r02c_dbg_pairs2.txt
