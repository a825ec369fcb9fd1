cd $GRAFT_REPO_ROOT
run() { for m in $2; do echo "=== type $1 mask $m"; SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $1 2>&1 | tail -6 | grep "demod ms" | cut -c1-60; done; }
run 0 "CCCFCC CCDFCC CDCFCC"
run 1 "CCCFCC CCDFCC"
run 2 "CCEECC CDEECC CCEFCC"
