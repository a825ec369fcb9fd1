cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
echo "=== host tests on 2 GPUs"
timeout 600 python -m pytest tests/test_host_cpp.py -x -q -m gpu 2>&1 | tail -4
echo "=== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scatter','single_source_wideband')}); print('e2e', d['e2e']['value'], d.get('e2e_wideband',{}).get('value'))"
