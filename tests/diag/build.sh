#!/bin/bash
# build the in-tree library from anywhere; non-zero exit on failure
set -e
cd "$(dirname "$0")/../.."
python -c "from cv_diffusion_model_b200 import build; print(build.build(verbose=False))"
