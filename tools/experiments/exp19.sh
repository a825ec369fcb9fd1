cd $GRAFT_REPO_ROOT
echo "=== full gpu test-suite"
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for c in 2 3 4 5; do
echo "=== bench config $c"
timeout 900 python bench.py --config $c --steps 10 --warmup 3 --seconds 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'], d['result'].get('auto'))"
done
