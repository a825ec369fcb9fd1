cd $GRAFT_REPO_ROOT
for a in "264 notma" "776 notma"; do
echo "=== RS41 $a"; timeout 60 python tools/dbg2.py 0 $a 2>&1 | tail -14
done
