#!/bin/bash
# Installs the unmodified reference under baseline/_ref (git-ignored; travels to the GPU box with gpurun).
# 1) the prescribed pip install (fails: the reference has no setup.py / pyproject.toml); 2) a plain copy of its
# first-party tree (everything except the vendored third-party/ submodules), which is what bench.py --impl reference
# reads its hyper-parameters from.
cd "$(dirname "$0")/.."
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference \
  > baseline/_ref_pip.log 2>&1 || echo "pip install failed (expected): $(tail -1 baseline/_ref_pip.log)"
mkdir -p baseline/_ref
if [ -d /root/reference ]; then
  (cd /root/reference && tar cf - --exclude=./third-party --exclude=./.git --exclude='*.pt.trace.json' .) | (cd baseline/_ref && tar xf -)
  echo "copied first-party reference tree: $(find baseline/_ref -type f | wc -l) files, $(du -sh baseline/_ref | cut -f1)"
fi
