cd $GRAFT_REPO_ROOT
echo "=== channelizer tests"
timeout 1200 python -m pytest tests/test_channelizer.py tests/test_host_cpp.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40
echo "=== chan bench"
timeout 300 python tools/chan_bench.py 1024 48 1 0 2>&1 | tail -2; timeout 300 python tools/chan_bench.py 1024 48 1 1 2>&1 | tail -2
