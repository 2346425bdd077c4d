cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/sweep.sh "PB2_X=0" "PB2_ORDER=colour" "PB2_CHUNK=14" "PB2_CHUNK=28" "PB2_CHUNK=8" "PB2_UNIT_PATCHES=1" "PB2_UNIT_PATCHES=1 PB2_CHUNK=14" "PB2_UNIT_PATCHES=8 PB2_CHUNK=14" "PB2_DEBUG_SCATTER=nostore" 2>&1
PB2_TIMING=1 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | grep "pb2 timing" | tail -1
