cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d.get('e2e')); print(d.get('scatter')); print(d.get('single_source_wideband'))"
grep -v "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r2_bench_2gpu.err | tail -5
