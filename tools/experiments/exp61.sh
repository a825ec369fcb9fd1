cd $GRAFT_REPO_ROOT
for c in 292 584; do echo "--- type 5, $c channels"; timeout 60 python tools/stalls.py 5 $c 2>&1 | tail -6 | cut -c1-100; done
