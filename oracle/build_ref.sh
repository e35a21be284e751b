#!/bin/sh
# oracle/build_ref.sh -- builds oracle/_ref/libocvref.so, the REFERENCE CPU arm of bench.py (cpu_baseline.kind = "reference").
#
# The reference (an OpenCV 3 fork) only configures through its own CMake tree (generated cvconfig.h / opencv_modules.hpp /
# OpenCL kernel sources, ~150 source files per module), which this project's rules exclude from the recipe; what IS allowed
# here is linking the UNMODIFIED reference CPU build that SURVEY.md Appendix A describes -- static libraries compiled from
# the reference's own sources (REFBUILD, default /tmp/refbuild; REFSRC = the source copy it was configured from, identical
# to /root/reference apart from nine CMake policy lines) -- behind the small C entry point oracle/refgen/ref_arm.cpp.
# Present only in the build container: on the GPU box the prebuilt .so is used as it travelled; without it bench.py falls
# back to the C port (cpu_baseline.kind = "port").  Output goes to oracle/_ref/ only (git-ignored).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
B=${REFBUILD:-/tmp/refbuild}
S=${REFSRC:-/tmp/refsrc}
[ -f "$B/lib/libopencv_octvr.a" ] || { echo "build_ref.sh: no reference CPU build at $B (skipped)"; exit 3; }
mkdir -p "$HERE/_ref"
INC="-I$B"
for m in core imgproc stitching features2d flann calib3d imgcodecs videoio highgui ml objdetect octvr; do INC="$INC -I$S/modules/$m/include"; done
/usr/bin/g++ -std=c++11 -O2 -w -fPIC -shared $INC "$HERE/refgen/ref_arm.cpp" -o "$HERE/_ref/libocvref.so" \
    -L"$B/lib" -Wl,--whole-archive -lopencv_octvr -Wl,--no-whole-archive -lopencv_stitching -lopencv_calib3d -lopencv_features2d -lopencv_flann \
    -lopencv_imgcodecs -lopencv_imgproc -lopencv_core -L"$B/3rdparty/lib" -llibjpeg -llibpng -lzlib -lpthread -ldl \
    -Wl,--exclude-libs,ALL
echo "built $HERE/_ref/libocvref.so"
