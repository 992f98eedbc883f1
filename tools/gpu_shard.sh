mkdir -p gpurun_out
for N in 2 8; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_shard$N.csv python tools/profile_shard.py cfg2 $N 2 > gpurun_out/launches_shard$N.log 2>&1
done
python tools/profile_shard.py cfg2 8 4 | tail -n 1
