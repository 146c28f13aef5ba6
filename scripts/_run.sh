bash scripts/gpu_check.sh
bash scripts/gpu_profile.sh
