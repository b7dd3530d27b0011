mkdir -p gpurun_out
cp simpledet_b200/libsimpledet_b200.so /tmp/orig.so
python -m pytest tests/test_roi_align_cl_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in NOPIPE PIPE_M4 PIPE_M3 PIPE_M2; do
  cp simpledet_b200/_ab/$v.so simpledet_b200/libsimpledet_b200.so
  for sh in target bench; do
    echo -n "$v "; python benchmarks/roi_align_sweep.py --shape $sh --path 4 --iters 30
  done
done 2>&1 | tee gpurun_out/r02_cl_ab10.txt
cp /tmp/orig.so simpledet_b200/libsimpledet_b200.so
