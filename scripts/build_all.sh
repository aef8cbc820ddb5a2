#!/bin/bash
# builds the product library and the diagnostics (timeline) library; fails loudly
set -e
cd "$(dirname "$0")/.."
python -m transkun_b200.build > /tmp/tkb_build.log 2>&1 || { grep -E "error" /tmp/tkb_build.log | head; echo "PRODUCT BUILD FAILED"; exit 1; }
python -c "from transkun_b200 import build; build.build_timeline()" > /tmp/tkb_build_tl.log 2>&1 || { grep -E "error" /tmp/tkb_build_tl.log | head; echo "TIMELINE BUILD FAILED"; exit 1; }
ls -la --time-style=+%H:%M:%S transkun_b200/csrc/*.so | awk '{print $6, $7}'
