mkdir -p gpurun_out
for w in infer retina_train mask_train dcn_softnms; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || echo "bench $w failed"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_infer_reference.json 2>/dev/null
for w in infer retina_train mask_train dcn_softnms; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
python benchmarks/roi_align_sweep.py --shape target --path 4 --iters 30
python benchmarks/roi_align_sweep.py --shape bench --path 4 --iters 30
python benchmarks/dcn_bench.py
python -c "
import json
for w in ['infer','retina_train','mask_train','dcn_softnms']:
    d=json.loads(open('gpurun_out/r02_bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['us_per_launch'], d.get('cpu_baseline',{}).get('value'), d['gpu_launches'], d.get('roofline_target',{}).get('frac'), d.get('roofline_target_nchw',{}).get('frac'))
"
