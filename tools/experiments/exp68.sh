cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_cli_dropin.py -x -q -m gpu -k "batch_runner" 2>&1 | tail -8
