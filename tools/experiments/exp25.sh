cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
for v in "SONDE_STREAM_PRIO=0" "SONDE_STREAM_PRIO=1" "SONDE_FRAME_SERIAL=1"; do
  for c in 2 3 5; do
    echo "=== cfg$c $v"
    env $v timeout 300 python bench.py --config $c --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
  done
done
