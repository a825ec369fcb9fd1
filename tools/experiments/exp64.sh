cd $GRAFT_REPO_ROOT
echo "=== full gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench default"; timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_s16']['value'], d['e2e_wideband']['value'], d['cpu_baseline']['value'], d['clocks'])"
