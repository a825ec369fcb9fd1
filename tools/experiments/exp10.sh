cd $GRAFT_REPO_ROOT
timeout 60 tools/ubench/ubench3
for a in "777" "48000" "1024 notma"; do
echo "=== RS41 $a"; timeout 60 python tools/dbg2.py 0 $a 2>&1 | tail -2
done
for m in 00CCCC 00DDCC CCCCCC CCDDCC; do
    echo "=== type 0 mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py 0 2>&1 | tail -6
done
