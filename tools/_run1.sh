mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "soft or nms or Nms" 2>&1 | tail -4
timeout 600 python bench.py --workload dcn_softnms --no-cpu-baseline | tail -1 | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_dcn_softnms.csv python bench.py --workload dcn_softnms --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
