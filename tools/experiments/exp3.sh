cd $GRAFT_REPO_ROOT
for m in CCCC DDCC; do
  echo "=== mask $m fm 1024"; SONDE_PW_MASK=$m timeout 120 python tools/dbg1.py 0 1024 fm 2>&1 | tail -8
done
echo "=== mask CCCC fm 48000"; SONDE_PW_MASK=CCCC timeout 120 python tools/dbg1.py 0 48000 fm 2>&1 | tail -8
echo "=== memcheck"; SONDE_PW_MASK=CCCC timeout 300 compute-sanitizer --tool memcheck python tools/dbg1.py 0 4096 fm 2>&1 | grep -v "^nsoft" | head -40
