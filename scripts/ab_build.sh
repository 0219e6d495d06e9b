#!/bin/bash
# Build the library of another git revision into ab_build/<name>/ (git-ignored, travels with gpurun) for same-box A/B runs:
#   scripts/ab_build.sh <rev> <name>;  B200W_LIB=$PWD/ab_build/<name>/libax_whisper.so python bench.py ...
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REV=$1; NAME=$2
SRC=$(mktemp -d)
git -C "$ROOT" archive "$REV" whisper.axera_b200/csrc include | tar -x -C "$SRC"
mkdir -p "$ROOT/ab_build/$NAME"
sed -i 's/^g++ .*whisper_srv.*$//' "$SRC/whisper.axera_b200/csrc/build.sh" || true
OBJ="$SRC/obj" OUT="$ROOT/ab_build/$NAME" bash "$SRC/whisper.axera_b200/csrc/build.sh" | tail -1
rm -rf "$SRC"
