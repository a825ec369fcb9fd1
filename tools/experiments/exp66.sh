cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match or bits_soft or seven or cfg4 or cfg3 or mixed" 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
for c in 4 3; do echo "=== cfg $c"; timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"; done
