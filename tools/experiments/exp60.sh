cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "afsk or seven or auto or iq_path or cfg5" 2>&1 | tail -3
for t in 5 6; do echo "--- type $t"; timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | cut -c1-100; done
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
echo "=== cfg 5"; timeout 300 python bench.py --config 5 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
echo "=== cfg5 timeline"; timeout 300 python tools/timeline.py 5 1024 3 2>&1 | tail -5 | cut -c1-150
