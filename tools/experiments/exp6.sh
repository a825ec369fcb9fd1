cd $GRAFT_REPO_ROOT
echo "=== RS41 777"; timeout 60 python tools/dbg2.py 0 777 2>&1 | tail -12
echo "=== RS41 780 notma"; timeout 60 python tools/dbg2.py 0 780 notma 2>&1 | tail -12
echo "=== RS41 1027"; timeout 60 python tools/dbg2.py 0 1027 2>&1 | tail -12
