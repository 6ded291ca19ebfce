# per-launch tensor-pipe activity, operand-path load, L2->SM and DRAM bytes of EVERY tensor-core convolution launch of one
# training pass (26 forward/dgrad + 13 wgrad) with the final code (metric list instead of --set full: 39 launches)
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cudnn"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpc__cycles_elapsed.avg.per_second
$CMD > $O/r02u_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:k_conv_umma -c 39 --csv --log-file $O/r02u_all_conv_metrics.csv $CMD > $O/r02u_all_conv.log 2>&1; echo rc=$?
