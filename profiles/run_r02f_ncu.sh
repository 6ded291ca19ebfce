# ncu evidence for the SM-pair kernels (run under gpurun after the plain command exited 0): launch list of the training step
# with DRAM bytes, and --set full of k_conv_umma_fwd4 / k_conv_umma_wgrad2 / k_conv_umma_wgrad_w3
TAG=${TAG:-r02f}
O=gpurun_out
set -x
python -m pytest tests/test_model_gpu.py -m gpu -x -q -s -k "trajectory" > $O/${TAG}_traj.log 2>&1; echo traj rc=$?; tail -5 $O/${TAG}_traj.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cudnn"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 \
    --csv --log-file $O/${TAG}_train_b8_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo list rc=$?
for K in k_conv_umma_fwd4 k_conv_umma_wgrad2; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 4 -f -o $O/${TAG}_full_$K $CMD > $O/${TAG}_full_$K.log 2>&1; echo $K rc=$?
done
python profiles/summarize_launches.py $O/${TAG}_train_b8_launches.csv $O/${TAG}_train_b8_launches_summary.md $O/${TAG}_train_b8_launches.json
head -24 $O/${TAG}_train_b8_launches_summary.md
