#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY: builds the host-side lane emulator used by tests/test_emulator.py.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
g++ -O2 -std=c++14 -fPIC -ffp-contract=off -shared -I"$here/../../include" -I"$here/../../longtr_b200/csrc" \
    -o "$here/libltr_emu.so" "$here/emu_viterbi.cpp"
