mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/r02_bench_infer.json 2> gpurun_out/r02_bench_infer.err; tail -c 300 gpurun_out/r02_bench_infer.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_infer.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_infer.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['us_per_launch'], d.get('cpu_baseline',{}).get('value'), d['gpu_launches'], d.get('roofline_target',{}).get('frac'), d.get('roofline_target_nchw',{}).get('frac'), d['clocks'])
"
