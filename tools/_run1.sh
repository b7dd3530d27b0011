python -m pytest tests/test_roi_align_cl_gpu.py -x -q -m gpu 2>&1 | tail -2
for sh in target bench; do python benchmarks/roi_align_sweep.py --shape $sh --path 4 --iters 30; done
