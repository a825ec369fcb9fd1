cd $GRAFT_REPO_ROOT
run() { for m in $2; do echo "=== type $1 mask $m"; SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $1 2>&1 | tail -6 | grep -v "PW last\|rounds"; done; }
run 0 "CCCFCC CCCCFCC 4CCCFCC CCCCCFCC"
run 2 "CCEECC CCCEECC"
