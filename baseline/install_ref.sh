#!/bin/bash
# Installs the UNMODIFIED reference into git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot,
# like the built libvdm.so) so that `bench.py --impl reference`, the cpu_baseline leg and the stock-GPU leg run the
# reference's own code.  Build container only: /root/reference does not exist on the GPU box.
#   1. the contract's pip install.  The reference's setup.py declares `py_modules=['improved_diffusion']` -- a package
#      directory named as a module -- so the wheel it builds holds metadata only (outcome recorded in DESIGN.md);
#   2. therefore the package directory the setup.py means to ship is placed next to that metadata, byte for byte.
# Nothing under baseline/_ref is tracked by git, and nothing in the product package imports it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/improved_diffusion" ] || { echo "no reference at $REF: keeping whatever baseline/_ref holds"; exit 0; }
rm -rf /tmp/vdm_refcopy "$ROOT/baseline/_ref"
cp -r "$REF" /tmp/vdm_refcopy
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" /tmp/vdm_refcopy || echo "pip install failed; continuing with the package copy"
if [ ! -f "$ROOT/baseline/_ref/improved_diffusion/unet.py" ]; then
  mkdir -p "$ROOT/baseline/_ref"
  cp -r "$REF/improved_diffusion" "$ROOT/baseline/_ref/improved_diffusion"
fi
rm -rf /tmp/vdm_refcopy
cmp "$REF/improved_diffusion/unet.py" "$ROOT/baseline/_ref/improved_diffusion/unet.py" && echo "baseline/_ref ready"
