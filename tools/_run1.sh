mkdir -p gpurun_out
python -m pytest tests/test_roi_align_cl_gpu.py -x -q -m gpu 2>&1 | tail -3
for sh in target bench; do for p in 4 3 1; do python benchmarks/roi_align_sweep.py --shape $sh --path $p --iters 30; done; done 2>&1 | tee gpurun_out/r02_cl_ab9.txt
python benchmarks/roi_align_sweep.py --shape target --path 4 --iters 30 --same-roi 1
