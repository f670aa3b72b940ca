#!/bin/bash
# Stage the UNMODIFIED reference package where the GPU box can see it:  /root/reference/statdepth  ->  oracle/_ref/statdepth
# (git-ignored, not gpurun-ignored: it travels with the working tree like the built .so files; /root/reference itself
# does not exist on the GPU box).  Nothing is patched; oracle/ref_shim.py supplies the two import shims (plotly stub,
# DataFrame.append) at run time.  Used ONLY by bench.py's CPU baseline (`reference_python`: the reference's own
# univariate band depth timed on the box's host cores) and by tests that skip without it.  Never imported by the product.
set -e
SRC="${STATDEPTH_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -d "$SRC/statdepth" ]; then
  echo "make_ref: $SRC/statdepth not present (GPU box?): keeping whatever is in oracle/_ref" >&2
  exit 0
fi
rm -rf "$HERE/_ref/statdepth"
mkdir -p "$HERE/_ref"
cp -r "$SRC/statdepth" "$HERE/_ref/statdepth"
find "$HERE/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "make_ref: staged $(find "$HERE/_ref/statdepth" -name '*.py' | wc -l) files under oracle/_ref/statdepth"
