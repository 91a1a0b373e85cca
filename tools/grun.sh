#!/bin/bash
# Build everything (product .so, oracle, oracle/_ref) here, then run a command on the B200 box.
# Usage: tools/grun.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
exec timeout $(( $1 + 1900 )) gpurun --timeout "$1" -- "$2"
