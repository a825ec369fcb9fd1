cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
echo "=== frame parity"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match or seven or pipelined_fetch" 2>&1 | tail -2
for v in "X=0" "SONDE_FRAME_SERIAL=1" "SONDE_FRAME_PERSIST=148 SONDE_FRAME_WORK=1 SONDE_FRAME_SKIP=3" "SONDE_FRAME_PERSIST=148 SONDE_FRAME_WORK=1 SONDE_FRAME_SKIP=2" "SONDE_FRAME_PERSIST=148 SONDE_FRAME_WORK=1 SONDE_FRAME_SKIP=0" "SONDE_FRAME_PERSIST=148 SONDE_FRAME_WORK=1 SONDE_FRAME_SKIP=1" "SONDE_FRAME_PERSIST=148 SONDE_FRAME_WORK=2 SONDE_FRAME_SKIP=2" "SONDE_FRAME_PERSIST=296 SONDE_FRAME_WORK=1 SONDE_FRAME_SKIP=3"; do
  echo "=== cfg2 $v"
  env $v timeout 300 python bench.py --config 2 --seconds 2 --steps 30 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
done
