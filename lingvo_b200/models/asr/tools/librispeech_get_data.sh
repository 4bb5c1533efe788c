#!/bin/bash
# Download + parameterise LibriSpeech. Usage: librispeech_get_data.sh ROOT_DIR [flags]
set -euo pipefail
ROOT="${1:?usage: $0 ROOT_DIR [flags]}"; shift || true
exec python -m lingvo_b200.models.asr.tools.librispeech_get_data --root="${ROOT}" "$@"
