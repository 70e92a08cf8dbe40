#!/usr/bin/env bash
# Build the reference's own executables (GetAlignmentStats, GetChromDepth, EstimateSVLoci, GenerateSVCandidates, ...)
# from the UNMODIFIED sources under $MANTA_REF (default /root/reference) into oracle/_ref/full/bin -- and, with
# --b200, a second GenerateSVCandidates whose refiner calls libmanta_b200.so (the patch of INTEGRATION.md section 5
# applied to a scratch copy of the GenerateSVCandidates directory under oracle/_ref/full/patched, never committed).
#
# The reference's cmake build is NOT run: htslib is built from the reference's redist tarball with its own Makefile,
# the boost libraries the executables link are compiled from the redist tarball's sources, the two generated headers
# (common/config.h, common/configBuildTimeInfo.h) are written here, every lib/*/ *.cpp is compiled once.
#
# TEST INFRASTRUCTURE ONLY (VCF-level parity, tests/test_vcf_demo.py): nothing here is part of the product.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MANTA_REF:-/root/reference}"
OUT="$HERE/_ref/full"
TP="$HERE/_ref"           # boost / htslib unpack dir shared with build_ref.sh
LIB="$REF/src/c++/lib"
JOBS="${JOBS:-8}"
WITH_B200=0
WITH_TRACE=0
for a in "$@"; do
  [ "$a" = "--b200" ] && WITH_B200=1
  [ "$a" = "--trace" ] && WITH_TRACE=1
done

if [ ! -d "$LIB" ]; then
  echo "build_ref_full.sh: reference tree not found at $REF (expected on the build container only)" >&2
  exit 3
fi
mkdir -p "$OUT/obj" "$OUT/bin" "$OUT/gen/common" "$OUT/demo"
if [ -x /usr/bin/g++ ]; then CXX=/usr/bin/g++; CC=/usr/bin/gcc; else CXX=g++; CC=gcc; fi

# --- htslib (static, gzip only: the demo inputs are BAM + plain fasta) -------------------------------------------
if [ ! -f "$OUT/htslib-1.9/libhts.a" ]; then
  echo "[full] building htslib"
  tar -xjf "$REF/redist/htslib-1.9.tar.bz2" -C "$OUT"
  (cd "$OUT/htslib-1.9" && ./configure CC=$CC --disable-bz2 --disable-lzma --disable-libcurl >/dev/null && make -j"$JOBS" lib-static >/dev/null 2>&1)
fi

# --- boost: headers + the compiled libraries the executables need --------------------------------------------------
BOOST="$TP/boost_1_58_0"
BLIBS="program_options filesystem system serialization date_time regex timer chrono"
for l in $BLIBS; do
  if [ ! -d "$BOOST/libs/$l/src" ]; then
    tar -xjf "$REF/redist/boost_1_58_0.tar.bz2" -C "$TP" boost_1_58_0/boost $(for k in $BLIBS; do echo boost_1_58_0/libs/$k/src; done)
    break
  fi
done

# --- generated headers ------------------------------------------------------------------------------------------
cat > "$OUT/gen/common/config.h" <<'EOF'
#pragma once
#include "configBuildTimeInfo.h"
#define STRINGIFY(x) #x
#define EXPAND(x) STRINGIFY(x)
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_UNISTD_H 1
#define HAVE_ZLIB 1
#define CXX_COMPILER_NAME "g++"
#define COMPILER_VERSION "oracle-build"
EOF
cat > "$OUT/gen/common/configBuildTimeInfo.h" <<'EOF'
#pragma once
#define BUILD_TIME "oracle-build"
#define WORKFLOW_VERSION "oracle-build"
EOF

CXXFLAGS="-std=c++11 -O3 -fomit-frame-pointer -w -DBOOST_SYSTEM_NO_DEPRECATED -DBOOST_ALL_NO_LIB \
  -I$LIB -I$OUT/gen -I$OUT/gen/common -I$BOOST -I$OUT/htslib-1.9 -I$REF/redist/ctpl-0.0.2"

compile_list() {   # stdin: "src obj" pairs
  xargs -P "$JOBS" -n 2 sh -c 'if [ ! -f "$1" ] || [ "$0" -nt "$1" ]; then mkdir -p "$(dirname "$1")"; '"$CXX $CXXFLAGS"' $EXTRA -c "$0" -o "$1" || exit 255; fi'
}

echo "[full] compiling boost libraries"
export EXTRA=""
for l in $BLIBS; do
  find "$BOOST/libs/$l/src" -name '*.cpp' | while read -r f; do
    case "$l/$(basename "$f" .cpp)" in
      serialization/*w*archive*|serialization/*wi*|serialization/*wo*|serialization/xml_w*|serialization/text_w*|serialization/xml_wgrammar|serialization/shared_ptr_helper) continue ;;
      regex/w*|regex/usinstances|regex/icu) continue ;;
    esac
    echo "$f $OUT/obj/boost/$l/$(basename "${f%.cpp}").o"
  done
done | compile_list
rm -f "$OUT/libboost_min.a"; ar rcs "$OUT/libboost_min.a" $(find "$OUT/obj/boost" -name '*.o')

echo "[full] compiling the reference libraries"
LIBDIRS="manta svgraph format alignment assembly options appstats htsapi common blt_util"
for d in $LIBDIRS; do
  for f in "$LIB/$d"/*.cpp; do echo "$f $OUT/obj/lib/$d/$(basename "${f%.cpp}").o"; done
done | compile_list
rm -f "$OUT/libmanta_all.a"
# link order of the reference (src/c++/lib/CMakeLists.txt PROJECT_PRIMARY_LIBRARY_DIRS) is irrelevant inside one archive
ar rcs "$OUT/libmanta_all.a" $(find "$OUT/obj/lib" -name '*.o')

APPS="GetAlignmentStats MergeAlignmentStats SummarizeAlignmentStats GetChromDepth EstimateSVLoci MergeSVLoci CheckSVLoci DumpSVLoci GenerateSVCandidates"
echo "[full] compiling the applications"
for a in $APPS; do
  for f in "$LIB/applications/$a"/*.cpp; do echo "$f $OUT/obj/app/$a/$(basename "${f%.cpp}").o"; done
  echo "$REF/src/c++/bin/$a.cpp $OUT/obj/app/$a/main_$a.o"
done | compile_list
LINK_LIBS="$OUT/libmanta_all.a $OUT/libboost_min.a $OUT/htslib-1.9/libhts.a -lz -lpthread -lrt"
for a in $APPS; do
  $CXX -o "$OUT/bin/$a" "$OUT"/obj/app/$a/*.o -Wl,--start-group $LINK_LIBS -Wl,--end-group
done
strip "$OUT"/bin/* 2>/dev/null || true   # the binaries travel to the GPU box with every gpurun snapshot
echo "[full] built $(ls "$OUT/bin" | tr '\n' ' ')"

# --- demo inputs (data, not sources) next to the binaries so that they travel to the GPU box --------------------
cp -u "$REF"/src/demo/data/*.bam "$REF"/src/demo/data/*.bai "$OUT/demo/"
# the fasta (286 MB, nearly all N) stays packed: run_demo.sh unpacks it where it runs
cp -u "$REF/src/demo/data/Homo_sapiens_assembly19.COST16011_region.fa.tar.bz2" "$REF/src/demo/data/Homo_sapiens_assembly19.COST16011_region.fa.fai" "$OUT/demo/"
cp -u "$REF/src/demo/expectedResults/somaticSV.vcf.gz" "$OUT/demo/expected_somaticSV.vcf.gz"

# --- the B200-backed GenerateSVCandidates -------------------------------------------------------------------------
if [ "$WITH_B200" = 1 ]; then
  GSC="$LIB/applications/GenerateSVCandidates"
  PD="$OUT/patched"
  rm -rf "$PD"; mkdir -p "$PD/applications/GenerateSVCandidates" "$PD/manta"
  # scratch copy of the GenerateSVCandidates directory (its sources include each other by bare file name, so the
  # patched header must sit next to all of them) + the one lib/manta file the patch touches; git-ignored like all of
  # oracle/_ref
  cp "$GSC"/*.cpp "$GSC"/*.hpp "$PD/applications/GenerateSVCandidates/"
  cp "$LIB/manta/SVCandidateAssembler.cpp" "$PD/manta/"
  (cd "$PD" && patch -p1 --quiet < "$HERE/../integration/manta_b200.patch")
  export EXTRA="-I$HERE/../include"
  { for f in "$PD"/applications/GenerateSVCandidates/*.cpp; do echo "$f $OUT/obj/b200/$(basename "${f%.cpp}").o"; done
    echo "$PD/manta/SVCandidateAssembler.cpp $OUT/obj/b200/lib_SVCandidateAssembler.o"
    echo "$REF/src/c++/bin/GenerateSVCandidates.cpp $OUT/obj/b200/main.o"
  } | compile_list
  # the patched assembler object must win over the archive member: objects first, archive after
  $CXX -o "$OUT/bin/GenerateSVCandidates_b200" "$OUT"/obj/b200/*.o -Wl,--start-group $LINK_LIBS -Wl,--end-group \
    -L"$HERE/../manta_b200" -lmanta_b200 -Wl,-rpath,'$ORIGIN/../../../../manta_b200'
  strip "$OUT/bin/GenerateSVCandidates_b200" 2>/dev/null || true
  echo "[full] built GenerateSVCandidates_b200"
fi

# --- the call-recording GenerateSVCandidates (stock algorithms + oracle/ref_trace.hpp at every hot-path call site) ----
if [ "$WITH_TRACE" = 1 ]; then
  GSC="$LIB/applications/GenerateSVCandidates"
  TD="$OUT/traced"
  rm -rf "$TD"; mkdir -p "$TD/applications/GenerateSVCandidates" "$TD/manta"
  cp "$GSC"/*.cpp "$GSC"/*.hpp "$TD/applications/GenerateSVCandidates/"
  cp "$LIB/manta/SVCandidateAssembler.cpp" "$TD/manta/"
  (cd "$TD" && patch -p1 --quiet < "$HERE/trace_calls.patch")
  export EXTRA="-I$HERE"
  { for f in "$TD"/applications/GenerateSVCandidates/*.cpp; do echo "$f $OUT/obj/traced/$(basename "${f%.cpp}").o"; done
    echo "$TD/manta/SVCandidateAssembler.cpp $OUT/obj/traced/lib_SVCandidateAssembler.o"
    echo "$REF/src/c++/bin/GenerateSVCandidates.cpp $OUT/obj/traced/main.o"
  } | compile_list
  $CXX -o "$OUT/bin/GenerateSVCandidates_trace" "$OUT"/obj/traced/*.o -Wl,--start-group $LINK_LIBS -Wl,--end-group
  strip "$OUT/bin/GenerateSVCandidates_trace" 2>/dev/null || true
  echo "[full] built GenerateSVCandidates_trace"
fi
