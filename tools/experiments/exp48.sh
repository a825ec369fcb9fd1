cd $GRAFT_REPO_ROOT
timeout 300 compute-sanitizer --tool memcheck python tools/dbg1.py 0 4096 fm 2>&1 | grep -v "Host Frame" | head -30
