cd $GRAFT_REPO_ROOT
for v in "" "SONDE_FRAME_SERIAL=1"; do
  echo "=== cfg2 $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'])"
done
echo "=== frames parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match or all_seven or auto" 2>&1 | tail -3
