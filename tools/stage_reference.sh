#!/bin/bash
# Stages the UNMODIFIED reference scripts where the drop-in test finds them on the GPU box (which has no /root/reference):
# baseline/_ref/ is git-ignored -- no reference source enters the history -- but travels with a gpurun snapshot.
set -e
src=${1:-/root/reference}
dst="$(dirname "$0")/../baseline/_ref"
mkdir -p "$dst"
cp "$src"/*.py "$dst"/
echo "staged $(ls "$dst"/*.py | wc -l) files from $src into $dst"
