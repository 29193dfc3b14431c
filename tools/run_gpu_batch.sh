cd $GRAFT_REPO_ROOT
bash tests/run_sanitizer.sh 2>&1 | tail -20
