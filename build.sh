#!/bin/bash
# Builds libbpgpu.so (sm_100a) in-tree; see __graft_entry__.build_lib (per-translation-unit objects, parallel).
set -e
cd "$(dirname "$0")"
python -c "import sys, __graft_entry__ as g; g.build_lib(sys.argv[1:])" "$@"
