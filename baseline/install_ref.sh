#!/bin/bash
# Populate baseline/_ref with the UNMODIFIED reference sources (git-ignored; travels with gpurun).
# `pip install /root/reference` is attempted first, as the build spec asks; the reference has neither
# setup.py nor pyproject.toml, so pip refuses and the sources are copied verbatim instead.
set -u
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference \
  > baseline/_ref/pip_install.log 2>&1 || echo "pip install failed (expected: no setup.py/pyproject.toml); copying sources"
mkdir -p baseline/_ref/src/utils
cp /root/reference/run.py baseline/_ref/src/
cp /root/reference/utils/*.py baseline/_ref/src/utils/
( cd /root/reference && sha256sum run.py utils/*.py ) > baseline/_ref/SHA256SUMS
( cd baseline/_ref/src && sha256sum -c ../SHA256SUMS ) | tail -3
# real MNIST for the GPU-box training runs (gz IDX files, 11 MB; git-ignored like the rest of baseline/_ref)
mkdir -p baseline/_ref/datasets/MNIST/raw
cp /root/reference/datasets/MNIST/raw/*.gz baseline/_ref/datasets/MNIST/raw/
