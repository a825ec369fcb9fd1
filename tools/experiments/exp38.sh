cd $GRAFT_REPO_ROOT
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "iq_path or mixed or ragged or pipelined or kernel_equals" 2>&1 | tail -3
for t in 0 1 2; do
  echo "=== type $t"
  timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | grep -v "PW last"
done
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frame_kernel_ms"], d["value"])'
echo "=== cfg 2"; timeout 300 python bench.py --config 2 --seconds 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P"
