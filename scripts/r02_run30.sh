set -x
cd $GRAFT_REPO_ROOT
bash scripts/ncu_capture.sh r02c
