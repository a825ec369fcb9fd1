cd $GRAFT_REPO_ROOT
echo "=== multibank"
timeout 600 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k multibank 2>&1 | tail -3
bash tools/sanitize.sh
