#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY: builds the host-side lane emulators used by tests/test_emulator*.py.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
inc="-I$here/../../include -I$here/../../longtr_b200/csrc"
g++ -O2 -std=c++14 -fPIC -ffp-contract=off -shared $inc -o "$here/libltr_emu.so" "$here/emu_viterbi.cpp"
g++ -O2 -std=c++14 -fPIC -ffp-contract=off -shared $inc -o "$here/libltr_emu_stutter.so" "$here/emu_stutter.cpp" \
    "$here/../../longtr_b200/csrc/host/host_types.cpp"
g++ -O2 -std=c++14 -fPIC -shared $inc -o "$here/libltr_emu_edit.so" "$here/emu_edit.cpp"
