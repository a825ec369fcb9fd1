cd $GRAFT_REPO_ROOT
for v in "" "SONDE_FRAME_SERIAL=1"; do
  echo "=== cfg2 $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'])"
done
