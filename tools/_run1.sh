mkdir -p gpurun_out
python -m pytest tests/test_dcn_gpu.py -x -q -m gpu 2>&1 | tail -3
python benchmarks/dcn_bench.py
ncu --set full --clock-control none --import-source on -k regex:deform_im2col_cl_kernel -c 1 -o gpurun_out/r02_dcn_cl -f python benchmarks/dcn_bench.py --iters 1 > /dev/null 2>&1
