cd $GRAFT_REPO_ROOT
for m in CCCCC CCDCC CCFCC CDFCC CFFCC CDDCC; do
  echo "=== cfg2 mask $m"
  SONDE_PW_MASK=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'])"
done
for m in CCCCC CDDCC CFFCC; do
  echo "=== cfg4 mask $m"
  SONDE_PW_MASK=$m timeout 300 python bench.py --config 4 --seconds 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frame_kernel_ms'])"
done
echo "=== cfg5 launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_cfg5_launches.csv python bench.py --config 5 --seconds 2 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -25 gpurun_out/r2_cfg5_launches.csv | cut -d, -f5,12- | cut -c1-150
echo "=== afsk tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "afsk or full_size or iq_path" 2>&1 | tail -5
