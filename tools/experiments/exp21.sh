cd $GRAFT_REPO_ROOT
echo "=== multibank test"
timeout 600 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k multibank 2>&1 | tail -30
echo "=== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}); print(d.get('scatter')); print(d.get('single_source_wideband'))"
