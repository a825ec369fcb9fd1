cd $GRAFT_REPO_ROOT
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== bench default"; (time timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err); tail -2 gpurun_out/r2_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e']); print(d.get('e2e_s16')); print(d.get('e2e_wideband')); print(d['cpu_baseline'])"
echo "=== reference arm"; (time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null); cut -c1-600 gpurun_out/r2_bench_reference_arm.json
