cd $GRAFT_REPO_ROOT
echo "=== parity"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" 2>&1 | tail -3
echo "=== stalls type 0"; timeout 60 python tools/stalls.py 0 2>&1 | tail -6 | grep -v "PW last\|rounds" | cut -c1-100
echo "=== stalls type 0, 1-D"; SONDE_NO_TMA2D=1 timeout 60 python tools/stalls.py 0 2>&1 | tail -6 | grep -v "PW last\|rounds" | cut -c1-100
echo "=== stalls type 2"; timeout 60 python tools/stalls.py 2 2>&1 | tail -6 | grep -v "PW last\|rounds" | cut -c1-100
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
echo "=== cfg 2"; timeout 300 python bench.py --config 2 --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
