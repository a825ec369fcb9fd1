cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "bch_error or frames_match or cfg4 or cfg5" 2>&1 | grep -v "^$" | tail -8
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
for c in 4; do echo "=== cfg $c"; timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"; done
