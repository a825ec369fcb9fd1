cd $GRAFT_REPO_ROOT
echo "=== full gpu test-suite"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for m in 00CCCC CCCCCC CCDDCC CCFFCC; do
    echo "=== type 0 mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py 0 2>&1 | tail -6
done
for t in 1 2; do
  for m in CCCCCC CCDDCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6
  done
done
