mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "retina or Retina or customop or dropin" 2>&1 | tail -4
timeout 600 python bench.py --workload retina_train > gpurun_out/r02_bench_retina_train.json 2> gpurun_out/r02_bench_retina_train.err || echo "bench failed"
tail -c 300 gpurun_out/r02_bench_retina_train.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_retina_train.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['us_per_launch'], d.get('cpu_baseline'))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_retina_train.csv python bench.py --workload retina_train --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
