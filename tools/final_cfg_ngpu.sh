cd $GRAFT_REPO_ROOT
N=$1; CFG=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --config $CFG --steps 20 --warmup 3 --no-scatter > gpurun_out/r2_bench_c${CFG}_${N}gpu.json 2> gpurun_out/r2_bench_c${CFG}_${N}gpu.err
python - <<PY
import json; d=json.load(open('gpurun_out/r2_bench_c${CFG}_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['workload'][:80])
print('result', {k:(round(v) if isinstance(v,(int,float)) else v) for k,v in d['result'].items() if k!='auto'}); print('auto', d['result'].get('auto'))
e=d.get('e2e',{}); print('e2e', e.get('value'), e.get('h2d_gbs_achieved'))
PY
grep -v "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r2_bench_c${CFG}_${N}gpu.err | tail -3
