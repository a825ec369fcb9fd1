cd $GRAFT_REPO_ROOT
time (timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err)
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d.get('e2e')); print(d.get('scatter')); print(d.get('single_source_wideband')); print(d.get('e2e_wideband'))"
tail -5 gpurun_out/bench_8gpu.err
