set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/gpu_diag.py 4096 1 148 1184
python scripts/gpu_diag.py 65536 1184
