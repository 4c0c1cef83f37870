#!/bin/bash
# Install the UNMODIFIED reference (mdtraj, /root/reference) into baseline/_ref/ (git-ignored, shipped by gpurun) so that
# tests/test_reference_integration.py can run the reference's own tests against mdtraj_b200.patch_mdtraj().
# Offline: the only things added are two build/run shims the image lacks -- a stub `versioneer` module next to the COPY
# of setup.py (upstream imports it for the version string only) and pip's vendored `pyparsing` (mdtraj/core/selection.py
# imports it; never touched on the RMSD path).  No reference source file is edited and nothing is copied into the repo.
set -e
cd "$(dirname "$0")/.."
SRC=/tmp/mdtraj_src_$$
rm -rf "$SRC" baseline/_ref
cp -r /root/reference "$SRC" && chmod -R u+w "$SRC"
cat > "$SRC/versioneer.py" <<'PY'
def get_version():
    return "0+reference"
def get_cmdclass():
    return {}
PY
CC=/usr/bin/gcc CXX=/usr/bin/g++ LDSHARED="/usr/bin/g++ -shared" python -m pip install --no-index --no-build-isolation \
    --no-deps --find-links /opt/wheelhouse --target baseline/_ref "$SRC" 2>&1 | tail -5
PYP=$(python -c "import pip, os; print(os.path.join(os.path.dirname(pip.__file__), '_vendor', 'pyparsing'))")
[ -d "$PYP" ] && cp -r "$PYP" baseline/_ref/pyparsing
rm -rf "$SRC"
PYTHONPATH=baseline/_ref python -c "import mdtraj as md, numpy as np; t = md.Trajectory(np.random.rand(5, 10, 3).astype('f'), None); print('reference import ok', md.__version__, md.rmsd(t, t, 0)[:2])"
