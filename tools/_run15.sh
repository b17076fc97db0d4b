set -x
python tools/time_frontback.py 2>&1 | tail -6
