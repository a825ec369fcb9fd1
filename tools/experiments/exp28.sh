cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
echo "=== cfg2 146 ch"; timeout 300 python tools/timeline.py 2 146 3 2>&1 | tail -6
echo "=== cfg5 nofork serial"; SONDE_NO_FORK=1 SONDE_FRAME_SERIAL=1 timeout 300 python tools/timeline.py 5 1024 2 2>&1 | tail -10
echo "=== cfg2 skip2"; SONDE_FRAME_SKIP=2 timeout 300 python tools/timeline.py 2 1024 5 2>&1 | tail -10
echo "=== bench cfg2 skip2"; SONDE_FRAME_SKIP=2 timeout 300 python bench.py --config 2 --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
echo "=== bench cfg2 skip0"; timeout 300 python bench.py --config 2 --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
