cd $GRAFT_REPO_ROOT
echo "=== cfg2"; timeout 300 python tools/timeline.py 2 1024 5 2>&1 | tail -14
echo "=== cfg5"; timeout 300 python tools/timeline.py 5 1024 4 2>&1 | tail -24
echo "=== cfg5 serial"; SONDE_FRAME_SERIAL=1 timeout 300 python tools/timeline.py 5 1024 4 2>&1 | tail -24
