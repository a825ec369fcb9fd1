cd $GRAFT_REPO_ROOT
echo "=== ubench2"; timeout 60 tools/ubench/ubench2
echo "=== RS41 1024 notma"; timeout 60 python tools/dbg1.py 0 1024 fm notma 2>&1 | tail -5
echo "=== RS41 776 tma"; timeout 60 python tools/dbg1.py 0 776 fm 2>&1 | tail -5
echo "=== RS41 772 tma"; timeout 60 python tools/dbg1.py 0 772 fm 2>&1 | tail -5
echo "=== RS41 777"; timeout 60 python tools/dbg1.py 0 777 fm 2>&1 | tail -5
echo "=== RS41 1033 (4 tiles + 9)"; timeout 60 python tools/dbg1.py 0 1033 fm 2>&1 | tail -5
echo "=== RS41 1027 "; timeout 60 python tools/dbg1.py 0 1027 fm 2>&1 | tail -5
