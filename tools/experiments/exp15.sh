cd $GRAFT_REPO_ROOT
echo "=== full gpu test-suite"
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
echo "=== bench config 2"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d.get(k) for k in ('value','ms_per_step','roofline','cpu_baseline')}); print(d['e2e']['value'], d.get('e2e_wideband',{}).get('value'))"
for c in 3 4 5; do
echo "=== bench config $c"
timeout 900 python bench.py --config $c --steps 6 --warmup 3 --seconds 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; tail -3 gpurun_out/bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c$c.json')); print({k:d.get(k) for k in ('value','ms_per_step','cpu_baseline','result')}); print(d['e2e']['value'])"
done
