mkdir -p gpurun_out
nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader; for d in /sys/bus/pci/devices/*/local_cpulist; do :; done; nproc
for w in infer retina_train; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w > gpurun_out/r02_bench_n2_$w.json 2> gpurun_out/r02_bench_n2_$w.err || echo "N=2 $w failed"
done
python -c "
import json
for w in ['infer','retina_train']:
    d=json.loads(open('gpurun_out/r02_bench_n2_%s.json'%w).read().strip().splitlines()[-1])
    print(w, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce'), d['config'].get('numa_binding'))
"
