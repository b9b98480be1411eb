set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python scripts/gpu_sweep.py 8192 1 8 > gpurun_out/sweep_r2d.log 2>&1; cat gpurun_out/sweep_r2d.log
python scripts/gpu_profile.py 8192 1184 > gpurun_out/phase_r2d.log 2>&1; cat gpurun_out/phase_r2d.log
