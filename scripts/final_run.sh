#!/bin/bash
# final evidence of a round: the driver's sequence, then the ncu passes (each only after its plain command exited 0)
bash scripts/gpu_check.sh
bash scripts/gpu_profile.sh
