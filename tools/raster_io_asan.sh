#!/bin/bash
# The raster I/O library's tests (codec round trips, libtiff cross-checks, boundless windows, 400 corrupted files) under
# AddressSanitizer + UndefinedBehaviorSanitizer.  Builds an instrumented libfz_rasterio.so in place, runs the CPU tests with
# the sanitizer runtimes preloaded into python, restores the release build.  Last run (final code of round 2): 40 passed, no report.
set -e
cd "$(dirname "$0")/.."
LIB=flair_for_aigle_b200/_native/libfz_rasterio.so
cp "$LIB" /tmp/libfz_rasterio.release.so
trap 'cp /tmp/libfz_rasterio.release.so "$LIB"' EXIT
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -shared -pthread -I include \
    flair_for_aigle_b200/csrc/host/raster_io.cpp -lz -o "$LIB"
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_rasterio.py tests/test_geotiff.py tests/test_zone_shards.py -x -q -p no:cacheprovider
