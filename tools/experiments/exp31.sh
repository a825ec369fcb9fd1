cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_channelizer.py tests/test_host_cpp.py -q -m gpu -s 2>&1 | grep -v "^$" | grep -v "^D=\|^channel " | tail -30
