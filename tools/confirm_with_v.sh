#!/bin/sh
# confirm_with_v.sh REF_DIR -- pins tests/golden/*.json (oracle outputs) to the real reference.
# Needs: a V compiler on PATH (`v`), a checkout of dy-tea/zpaq-v at REF_DIR, python3 with numpy.
# One command; exit 0 = every compressed block of the golden set (levels 0-5 x 9 inputs) and the tiny
# journaling archive are byte-identical (length + SHA-1 / hex) between the oracle and the V reference.
set -e
REF=${1:?usage: tools/confirm_with_v.sh /path/to/zpaq-v}
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v v >/dev/null || { echo "no V compiler on PATH" >&2; exit 2; }
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
python3 "$HERE/tests/golden/make_golden.py" --write-inputs "$WORK/inputs"
# a scratch module root next to the unmodified reference sources: v.mod + zpaq/ (symlink) + our main
mkdir -p "$WORK/root/golden"
cp "$REF/v.mod" "$WORK/root/v.mod"
ln -s "$(cd "$REF" && pwd)/zpaq" "$WORK/root/zpaq"
cp "$HERE/tools/ref_golden.v" "$WORK/root/golden/main.v"
(cd "$WORK/root" && v -o "$WORK/ref_golden" golden/)
"$WORK/ref_golden" "$WORK/inputs" > "$WORK/ref.txt"
python3 "$HERE/tests/golden/make_golden.py" --compare "$WORK/ref.txt"
