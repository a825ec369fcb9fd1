cd $GRAFT_REPO_ROOT
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
e=d.get('e2e',{}); print('e2e', e.get('value'), e.get('h2d_gbs_achieved'), e.get('h2d_link_gbs_measured'), e.get('h2d_link_gbs_all_ranks_at_once'))
print('s16', d.get('e2e_s16',{}).get('value'), 'wide', d.get('e2e_wideband',{}).get('value'))
s=d.get('scatter',{}); print('scatter', s.get('ms_per_step_with_scatter'), s.get('value_with_scatter'), s.get('rank0_egress_gbs'))
w=d.get('single_source_wideband',{}); print('ssw', w.get('ms_per_step'), w.get('value'), w.get('efficiency_vs_resident_shards'), w.get('rank0_egress_gbs'))
PY
grep -v "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r2_bench_${N}gpu.err | tail -4
