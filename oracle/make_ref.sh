#!/bin/sh
# oracle/make_ref.sh - stage the UNMODIFIED hot-path modules of the reference under oracle/_ref/.
#
# TEST / BENCHMARK INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (the reference's sources never enter
# this repository's history) but NOT gpurun-ignored, so the staged copy travels to the GPU box, where
# /root/reference does not exist.  It is used by
#   * bench.py --impl reference  (the reference's own run_pmpet / abcd_execute / streamrouting timed on the
#     box's host cores, cpu_baseline.kind = "reference"),
#   * bench.py's cpu_baseline leg,
#   * oracle/ref_loader.py as the fall-back location of the reference modules.
# Nothing under xanthos_b200/ may import it.
#
# Only the modules of the hot path are staged: they need numpy, scipy and joblib and
# nothing else, so no stubs for configobj / matplotlib are required.  The package __init__ files are created
# empty (the reference's xanthos/__init__.py imports the whole model, which needs configobj).
set -eu
SRC="${XANTHOS_REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/xanthos" ]; then
    echo "make_ref: no reference tree at $SRC (keeping whatever is in $DST)" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/xanthos"
: > "$DST/xanthos/__init__.py"
for f in pet/penman_monteith.py pet/hargreaves_samani.py pet/thornthwaite.py pet/hargreaves.py \
         runoff/abcd.py runoff/gwam.py routing/mrtm.py calibrate/calibrate_abcd.py utils/general.py; do
    d="$DST/xanthos/$(dirname "$f")"
    mkdir -p "$d"
    [ -f "$d/__init__.py" ] || : > "$d/__init__.py"
    cp "$SRC/xanthos/$f" "$d/"
done
for f in LICENSE DISCLAIMER; do [ -f "$SRC/$f" ] && cp "$SRC/$f" "$DST/"; done
( cd "$SRC/xanthos" && sha256sum pet/penman_monteith.py pet/hargreaves_samani.py pet/thornthwaite.py pet/hargreaves.py \
    runoff/abcd.py runoff/gwam.py routing/mrtm.py calibrate/calibrate_abcd.py utils/general.py ) > "$DST/SHA256SUMS"
echo "make_ref: staged $(wc -l < "$DST/SHA256SUMS") reference modules under $DST"
