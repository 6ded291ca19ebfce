# 1-GPU evidence of round 2 (run under gpurun): bench lines, ncu launch list (time + DRAM bytes per launch) and
# `--set full` captures of the dominant tensor-core kernels.  TAG selects the output prefix.
TAG=${TAG:-r02a}
O=gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo ref rc=$?
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cudnn"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 \
    --csv --log-file $O/${TAG}_train_b8_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo list rc=$?
for K in k_conv_umma_fwd2 k_conv_umma_fwd3 k_conv_umma_wgrad_w3; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -f -o $O/${TAG}_full_$K $CMD > $O/${TAG}_full_$K.log 2>&1; echo $K rc=$?
done
python profiles/summarize_launches.py $O/${TAG}_train_b8_launches.csv $O/${TAG}_train_b8_launches_summary.md $O/${TAG}_train_b8_launches.json
head -30 $O/${TAG}_train_b8_launches_summary.md
