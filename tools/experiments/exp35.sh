cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
for c in 3 5; do echo "=== cfg $c"; timeout 300 python bench.py --config $c --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"; done
echo "=== cfg5 timeline"; timeout 300 python tools/timeline.py 5 1024 3 2>&1 | tail -6
echo "=== afsk stalls"; timeout 100 python tools/stalls.py 5 2>&1 | tail -12
