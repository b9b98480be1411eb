set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/gpu_diag.py 4096 1 148 1036
python scripts/gpu_profile.py 8192 1
python scripts/gpu_diag.py 65536 1036
