cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --seconds 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d.get('scatter')); print(d.get('single_source_wideband'))"
tail -5 gpurun_out/bench_2gpu.err
