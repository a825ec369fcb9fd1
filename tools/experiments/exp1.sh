set -x
cd $GRAFT_REPO_ROOT
for t in 0 1 2; do
  for lay in 0 1 2; do
    echo "=== type $t layout $lay"
    SONDE_LAYOUT=$lay timeout 300 python tools/stalls.py $t 2>&1 | tail -7
  done
done
echo "=== ubench"
timeout 120 tools/ubench/ubench
