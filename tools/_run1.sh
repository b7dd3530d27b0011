python -m pytest tests -x -q -m gpu -k "backward or bwd or grad or autograd or roi_align" 2>&1 | tail -4
python benchmarks/roi_align_sweep.py --shape target --backward 1 --iters 20
python benchmarks/roi_align_sweep.py --shape train --backward 1 --iters 20
python benchmarks/roi_align_sweep.py --shape mask --backward 1 --iters 20
timeout 600 python bench.py --workload mask_train --no-cpu-baseline | tail -1 | cut -c1-200
