timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/gputests_r2m.log 2>&1; echo tests rc=$?; tail -n 2 gpurun_out/gputests_r2m.log
CMD="python bench.py --steps 6 --warmup 3 --no-small --no-cpu-baseline --no-ppo --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2m.csv $CMD > gpurun_out/ncu_launches_r2m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:quadx_ -s 6 -c 4 -f -o gpurun_out/k1_r2m $CMD > gpurun_out/ncu_full_r2m.log 2>&1
tail -n 1 gpurun_out/ncu_full_r2m.log
