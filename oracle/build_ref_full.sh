#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *whole* unmodified reference pipeline (everything poppy::morph<Twriter>() runs:
# extractor, matcher, transformer, procrustes, face, gabor_filter, blur_margin and the frame loop of
# src/poppy.hpp:46-248) as a harness binary, for BASELINE.json configs 1-3.
#
# What this does (nothing is copied into the repo; sources are compiled where they lie under /root/reference):
#   1. vendored OpenCV 4.6.0 core, imgproc, imgcodecs, features2d, flann, video, calib3d, photo, objdetect + contrib
#      face (vendored libpng / libjpeg-turbo) -> static libs in a scratch dir
#   2. reference src/{algo,util,draw,settings,face,extractor,matcher,transformer,procrustes,terminal}.cpp (-D_WASM:
#      drops HighGUI only) + oracle/ref_full_harness.cpp (ours: replaces SDL image loading by cv::imread and the boost
#      CLI by poppy::init, exactly the two things SURVEY B.2 replaces)        -> oracle/_ref/poppy_ref_full
#   3. the same + integration/algo_b200.cpp (the drop-in stub of INTEGRATION.md compiled against the reference's own
#      src/algo.hpp) linked with poppy_b200/libpoppy_cuda.so                  -> oracle/_ref/poppy_dropin
#      src/algo.cpp is compiled with -Dmorph_images=morph_images_reference (a command-line rename, the file is
#      untouched), so that the harness can route poppy::morph<Sink>()'s call at src/poppy.hpp:215 to either body.
# Outputs go only to oracle/_ref/ (git-ignored, but it travels to the GPU box). The reference's Makefile is NOT run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${POPPY_REFERENCE:-/root/reference}"
OCV="$REF/third/opencv-4.6.0"
CONTRIB="$REF/third/opencv_contrib-4.x/modules"
SCRATCH="${POPPY_REF_FULL_SCRATCH:-/tmp/poppy_ref_full}"
OUT="$HERE/_ref"
if [ ! -d "$OCV" ]; then echo "reference tree not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p "$SCRATCH/ocv" "$OUT"
if [ ! -f "$SCRATCH/ocv/lib/libopencv_face.a" ]; then
  cmake -G Ninja -S "$OCV" -B "$SCRATCH/ocv" -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DCMAKE_POSITION_INDEPENDENT_CODE=ON -DOPENCV_EXTRA_MODULES_PATH="$CONTRIB" \
    -DBUILD_LIST=core,imgproc,imgcodecs,features2d,flann,video,calib3d,photo,objdetect,face \
    -DBUILD_SHARED_LIBS=OFF -DWITH_IPP=OFF -DWITH_ITT=OFF -DWITH_OPENCL=OFF -DWITH_CUDA=OFF \
    -DWITH_TBB=OFF -DWITH_OPENMP=OFF -DWITH_EIGEN=OFF -DWITH_LAPACK=OFF -DWITH_PROTOBUF=OFF -DWITH_ADE=OFF -DWITH_QUIRC=OFF \
    -DWITH_PNG=ON -DBUILD_PNG=ON -DWITH_JPEG=ON -DBUILD_JPEG=ON -DWITH_TIFF=OFF -DWITH_WEBP=OFF -DWITH_OPENJPEG=OFF \
    -DWITH_JASPER=OFF -DWITH_OPENEXR=OFF \
    -DWITH_FFMPEG=OFF -DWITH_GSTREAMER=OFF -DWITH_V4L=OFF -DWITH_GTK=OFF -DWITH_QT=OFF -DWITH_1394=OFF -DBUILD_ZLIB=ON \
    -DBUILD_TESTS=OFF -DBUILD_PERF_TESTS=OFF -DBUILD_EXAMPLES=OFF -DBUILD_opencv_apps=OFF -DBUILD_JAVA=OFF \
    -DBUILD_opencv_python2=OFF -DBUILD_opencv_python3=OFF > "$SCRATCH/cmake.log" 2>&1
  ninja -C "$SCRATCH/ocv" -j"$(nproc)" > "$SCRATCH/ninja.log" 2>&1
fi
if [ "${1:-}" = "--opencv-only" ]; then echo "opencv built in $SCRATCH/ocv"; exit 0; fi

INC=(-I"$REF/src" -I"$SCRATCH/ocv" -I"$OCV/include" -I"$ROOT/include")
for m in core imgproc features2d flann video videoio highgui imgcodecs calib3d photo objdetect ml dnn stitching; do
  INC+=(-I"$OCV/modules/$m/include")
done
INC+=(-I"$CONTRIB/face/include")
CXX=(g++ -std=c++20 -O3 -D_WASM -w -pthread "${INC[@]}")
OBJ="$SCRATCH/obj"
mkdir -p "$OBJ"
# reference translation units, compiled where they lie (objects cached in the scratch dir)
pids=()
for tu in util draw settings face extractor matcher transformer procrustes terminal; do
  if [ ! -f "$OBJ/$tu.o" ] || [ "$REF/src/$tu.cpp" -nt "$OBJ/$tu.o" ]; then
    "${CXX[@]}" -c "$REF/src/$tu.cpp" -o "$OBJ/$tu.o" & pids+=($!)
  fi
done
if [ ! -f "$OBJ/algo_ref.o" ]; then
  "${CXX[@]}" -Dmorph_images=morph_images_reference -c "$REF/src/algo.cpp" -o "$OBJ/algo_ref.o" & pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
LIBS=(-L"$SCRATCH/ocv/lib" -L"$SCRATCH/ocv/3rdparty/lib" -lopencv_face -lopencv_objdetect -lopencv_photo -lopencv_calib3d
      -lopencv_features2d -lopencv_flann -lopencv_video -lopencv_imgcodecs -lopencv_imgproc -lopencv_core
      -llibpng -llibjpeg-turbo -lzlib -ldl -lpthread)
REFOBJS=("$OBJ"/{util,draw,settings,face,extractor,matcher,transformer,procrustes,terminal,algo_ref}.o)
"${CXX[@]}" "$HERE/ref_full_harness.cpp" "${REFOBJS[@]}" "${LIBS[@]}" -o "$OUT/poppy_ref_full"
echo "built $OUT/poppy_ref_full"
if [ -f "$ROOT/poppy_b200/libpoppy_cuda.so" ] && [ -f "$ROOT/integration/algo_b200.cpp" ]; then
  "${CXX[@]}" -DPOPPY_WITH_B200 -Dmorph_images=morph_images_b200 -c "$ROOT/integration/algo_b200.cpp" -o "$OBJ/algo_b200.o"
  # the conditioning drop-in: src/util.cpp once more with its two functions renamed (the file is untouched), and the stub
  if [ ! -f "$OBJ/util_ref.o" ]; then
    "${CXX[@]}" -Dblur_margin=blur_margin_reference -Dgabor_filter=gabor_filter_reference -c "$REF/src/util.cpp" -o "$OBJ/util_ref.o"
  fi
  "${CXX[@]}" -Dblur_margin=blur_margin_b200 -Dgabor_filter=gabor_filter_b200 -c "$ROOT/integration/util_b200.cpp" -o "$OBJ/util_b200.o"
  REFOBJS=("$OBJ"/{util_ref,draw,settings,face,extractor,matcher,transformer,procrustes,terminal,algo_ref}.o)
  "${CXX[@]}" -DPOPPY_WITH_B200 "$HERE/ref_full_harness.cpp" "$OBJ/algo_b200.o" "$OBJ/util_b200.o" "${REFOBJS[@]}" "${LIBS[@]}" \
      -L"$ROOT/poppy_b200" -lpoppy_cuda -Wl,-rpath,'$ORIGIN/../../poppy_b200' -o "$OUT/poppy_dropin"
  echo "built $OUT/poppy_dropin"
fi
