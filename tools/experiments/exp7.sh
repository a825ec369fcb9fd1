cd $GRAFT_REPO_ROOT
for a in "780" "896 notma" "896" "1032" "1032 notma" "776 notma" "520 notma" "264 notma" "200 notma"; do
echo "=== RS41 $a"; timeout 60 python tools/dbg2.py 0 $a 2>&1 | tail -4
done
