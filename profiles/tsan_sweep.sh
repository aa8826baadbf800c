#!/bin/bash
# ThreadSanitizer sweep of the kernel emulator (CPU only): every block of every kernel runs as T real threads with real
# barriers (SPIM_EMU_THREADS, csrc/hd.h) on a -fsanitize=thread build of the emulator, under the emulator test suites.
#   bash profiles/tsan_sweep.sh [threads=3]
# Prints the number of data-race reports (expected: 0) and restores the ordinary emulator build afterwards.
set -u
cd "$(dirname "$0")/.."
T=${1:-3}
TSAN=$(gcc -print-file-name=libtsan.so)
mkdir -p tests/emu
python spim_registration_b200/build.py > /dev/null || exit 1      # up to date, so that no test re-runs nvcc under the preloaded sanitizer
g++ -std=c++17 -O1 -g -fPIC -shared -fsanitize=thread -x c++ -DSPIM_HOST_EMU spim_registration_b200/csrc/spim_b200.cu -o tests/emu/libspim_emu.so || exit 1
touch tests/emu/libspim_emu.so
rm -f /tmp/spim_tsan_sweep.*
LD_PRELOAD=$TSAN TSAN_OPTIONS="exitcode=0 report_signal_unsafe=0 log_path=/tmp/spim_tsan_sweep" SPIM_EMU_THREADS=$T \
    python -m pytest tests/test_fusion_emulator.py tests/test_emulator.py tests/test_golden.py tests/test_bricks_p2p_threads.py -q -p no:cacheprovider 2>&1 | tail -3
echo "data-race reports: $(cat /tmp/spim_tsan_sweep.* 2>/dev/null | grep -c 'WARNING: ThreadSanitizer')"
rm -f tests/emu/libspim_emu.so
python -c "import __graft_entry__ as g; g.build_emulator(force=True)"
