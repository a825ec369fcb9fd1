cd $GRAFT_REPO_ROOT
echo "=== full gpu test-suite"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== bench cfg 2 3 5"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
for c in 2 3 5; do timeout 300 python bench.py --config $c --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"; done
echo "=== ncu K1 (production mask)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:demod_pipe -s 2 -c 1 -o gpurun_out/r2_k1 -f python tools/ncu_rs41.py > gpurun_out/r2_k1.log 2>&1; tail -2 gpurun_out/r2_k1.log
echo "=== ncu frame kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/r2_frame -f python tools/ncu_rs41.py > gpurun_out/r2_frame.log 2>&1; tail -2 gpurun_out/r2_frame.log
