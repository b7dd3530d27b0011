python -m pytest tests -x -q -m gpu -k "proposal_target or ProposalTarget or mask or Mask or config" 2>&1 | tail -3
timeout 600 python bench.py --workload mask_train --no-cpu-baseline | tail -1 | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_mask_train.csv python bench.py --workload mask_train --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
