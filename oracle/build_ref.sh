#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *unmodified* reference morph path as a shared library.
#
# What this does (nothing is copied into the repo; sources are compiled where they lie):
#   1. vendored OpenCV 4.6.0 core+imgproc (reference tree third/opencv-4.6.0) -> static libs, in a scratch dir
#   2. reference src/{algo,util,draw,settings}.cpp (-D_WASM -D_NO_FACE_DETECT: drops HighGUI / face linkage only)
#      + oracle/ref_shim.cpp (ours: a C-ABI veneer)  ->  oracle/_ref/libpoppy_ref.so
# Outputs go only to oracle/_ref/ (git-ignored, but it travels to the GPU box).
# The reference's own Makefile is NOT run; OpenCV is the arithmetic spec of the path.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${POPPY_REFERENCE:-/root/reference}"
OCV="$REF/third/opencv-4.6.0"
SCRATCH="${POPPY_REF_SCRATCH:-/tmp/poppy_ref_build}"
OUT="$HERE/_ref"
if [ ! -d "$OCV" ]; then echo "reference tree not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p "$SCRATCH/ocv" "$OUT"
if [ ! -f "$SCRATCH/ocv/lib/libopencv_imgproc.a" ]; then
  cmake -G Ninja -S "$OCV" -B "$SCRATCH/ocv" -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DCMAKE_POSITION_INDEPENDENT_CODE=ON \
    -DBUILD_LIST=core,imgproc -DBUILD_SHARED_LIBS=OFF -DWITH_IPP=OFF -DWITH_ITT=OFF -DWITH_OPENCL=OFF -DWITH_CUDA=OFF \
    -DWITH_TBB=OFF -DWITH_OPENMP=OFF -DWITH_EIGEN=OFF -DWITH_LAPACK=OFF -DWITH_PROTOBUF=OFF -DWITH_ADE=OFF -DWITH_QUIRC=OFF \
    -DWITH_PNG=OFF -DWITH_JPEG=OFF -DWITH_TIFF=OFF -DWITH_WEBP=OFF -DWITH_OPENJPEG=OFF -DWITH_JASPER=OFF -DWITH_OPENEXR=OFF \
    -DWITH_FFMPEG=OFF -DWITH_GSTREAMER=OFF -DWITH_V4L=OFF -DWITH_GTK=OFF -DWITH_QT=OFF -DWITH_1394=OFF -DBUILD_ZLIB=ON \
    -DBUILD_TESTS=OFF -DBUILD_PERF_TESTS=OFF -DBUILD_EXAMPLES=OFF -DBUILD_opencv_apps=OFF -DBUILD_JAVA=OFF \
    -DBUILD_opencv_python2=OFF -DBUILD_opencv_python3=OFF > "$SCRATCH/cmake.log" 2>&1
  ninja -C "$SCRATCH/ocv" -j"$(nproc)" > "$SCRATCH/ninja.log" 2>&1
fi
INC=(-I"$REF/src" -I"$SCRATCH/ocv" -I"$OCV/include")
for m in core imgproc features2d flann video videoio highgui imgcodecs calib3d photo objdetect ml dnn stitching; do
  INC+=(-I"$OCV/modules/$m/include")
done
g++ -std=c++20 -O3 -fPIC -shared -D_WASM -D_NO_FACE_DETECT -w -pthread "${INC[@]}" \
    "$HERE/ref_shim.cpp" "$REF/src/algo.cpp" "$REF/src/util.cpp" "$REF/src/draw.cpp" "$REF/src/settings.cpp" \
    -L"$SCRATCH/ocv/lib" -L"$SCRATCH/ocv/3rdparty/lib" \
    -Wl,--whole-archive -Wl,--no-whole-archive -lopencv_imgproc -lopencv_core -lzlib -ldl -lpthread \
    -Wl,--exclude-libs,ALL -o "$OUT/libpoppy_ref.so"
echo "built $OUT/libpoppy_ref.so"
