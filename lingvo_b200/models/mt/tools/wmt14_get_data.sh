#!/bin/bash
# One-shot data preparation for wmt14 (download → unpack → tokenize → wpm-encode).
# Usage: wmt14_get_data.sh /path/to/root [extra flags of wmt_get_data.py]
set -euo pipefail
ROOT="${1:?usage: $0 ROOT_DIR [flags]}"; shift || true
exec python -m lingvo_b200.models.mt.tools.wmt_get_data --dataset=wmt14 --root="${ROOT}" "$@"
