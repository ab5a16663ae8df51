timeout 900 python -m pytest tests -m gpu -q -x -k "sfm or rollout or integrate" 2>&1 | tail -4
python scripts/prof_sfm_rollout.py --scenes 1; python scripts/prof_sfm_rollout.py --scenes 64; python scripts/prof_sfm_rollout.py --scenes 4096 --frames 100
