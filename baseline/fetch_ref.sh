#!/bin/sh
# Vendors the UNMODIFIED reference files of the hot path into baseline/_ref/ (git-ignored; it travels
# to the GPU box with the gpurun snapshot) so that `bench.py --impl reference` and the `cpu_baseline`
# leg time the reference's own classes instead of the oracle port.
#
# The reference is a plain source tree (no setup.py / pyproject.toml), so
#   python -m pip install --no-index --no-build-isolation --target baseline/_ref /root/reference
# has nothing to install ("neither 'setup.py' nor 'pyproject.toml' found"); the path needs exactly
# three of its files, which import only torch / torchvision / numpy / scipy:
#   simpleAICV/detection/losses.py          RetinaLoss, FCOSLoss, IoUMethod
#   simpleAICV/detection/decode.py          RetinaDecoder, FCOSDecoder, DecodeMethod, DetNMSMethod
#   simpleAICV/detection/models/anchor.py   RetinaAnchors, FCOSPositions
# simpleAICV/detection/models/__init__.py imports every backbone of the repo; an EMPTY stub stands in
# for it (the two package __init__.py above it are empty in the reference as well).  Nothing is
# edited: losses.py's unused `from traitlets import Instance` is satisfied by a shim module at load
# time (baseline/refarm.py), exactly like tests/refload.py does.
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/simpleAICV/detection/losses.py" ]; then
    echo "fetch_ref: no reference checkout at $SRC (keeping $DST as it is)" >&2
    exit 0
fi
mkdir -p "$DST/simpleAICV/detection/models"
: > "$DST/simpleAICV/__init__.py"
: > "$DST/simpleAICV/detection/__init__.py"
: > "$DST/simpleAICV/detection/models/__init__.py"
cp "$SRC/simpleAICV/detection/losses.py" "$DST/simpleAICV/detection/losses.py"
cp "$SRC/simpleAICV/detection/decode.py" "$DST/simpleAICV/detection/decode.py"
cp "$SRC/simpleAICV/detection/models/anchor.py" "$DST/simpleAICV/detection/models/anchor.py"
( cd "$SRC" && sha256sum simpleAICV/detection/losses.py simpleAICV/detection/decode.py \
      simpleAICV/detection/models/anchor.py ) > "$DST/SHA256SUMS"
echo "fetch_ref: vendored 3 reference files into $DST"
