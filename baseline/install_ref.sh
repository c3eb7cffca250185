#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box).
# /root/reference is read-only and setup.py writes an egg-info next to itself, hence the copy;
# `future` (install_requires) is not in the offline wheelhouse, hence --no-deps -- the one name the
# reference imports from it is provided by baseline/shims/past/builtins.py.
set -e
cd "$(dirname "$0")/.."
[ -d /root/reference ] || { echo "no /root/reference here"; exit 0; }
rm -rf /tmp/_nb_refcopy baseline/_ref && cp -r /root/reference /tmp/_nb_refcopy
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/_nb_refcopy
diff -r -x __pycache__ /root/reference/numbskull baseline/_ref/numbskull && echo "baseline/_ref/numbskull identical to the reference"
