# ncu launch list (time + DRAM bytes per launch) of the training step with the current kernel sources; feeds roofline.traffic
TAG=${TAG:-r02n}
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cudnn"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 \
    --csv --log-file $O/${TAG}_train_b8_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo list rc=$?
python profiles/summarize_launches.py $O/${TAG}_train_b8_launches.csv $O/${TAG}_train_b8_launches_summary.md $O/${TAG}_train_b8_launches.json
head -16 $O/${TAG}_train_b8_launches_summary.md
