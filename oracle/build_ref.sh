#!/usr/bin/env bash
# Build oracle/_ref/libmanta_ref.so from the UNMODIFIED reference sources where they lie under
# $MANTA_REF (default /root/reference).  Nothing is copied into the repo: boost / htslib headers are
# unpacked from the reference's own redist tarballs into oracle/_ref/ (git-ignored), objects and the
# shared library go to oracle/_ref/ as well.  The reference's own build system (cmake) is NOT run;
# the handful of translation units the hot path needs are compiled directly.
#
# TEST INFRASTRUCTURE ONLY: the product (libmanta_b200.so) never links or loads this library.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MANTA_REF:-/root/reference}"
OUT="$HERE/_ref"
LIB="$REF/src/c++/lib"
GSC="$LIB/applications/GenerateSVCandidates"

if [ ! -d "$LIB" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the build container only)" >&2
  exit 3
fi
mkdir -p "$OUT/obj"

# --- third-party headers from the reference's redist tarballs ---------------------------------
if [ ! -d "$OUT/boost_1_58_0/boost" ]; then
  echo "[build_ref] unpacking boost headers"
  tar -xjf "$REF/redist/boost_1_58_0.tar.bz2" -C "$OUT" boost_1_58_0/boost boost_1_58_0/libs/date_time/src
fi
if [ ! -d "$OUT/boost_1_58_0/libs/serialization/src" ]; then
  # SVLocusSet.cpp instantiates boost::archive (graph save/load): the serialization library itself is compiled too
  echo "[build_ref] unpacking boost.serialization sources"
  tar -xjf "$REF/redist/boost_1_58_0.tar.bz2" -C "$OUT" boost_1_58_0/libs/serialization/src
fi
if [ ! -d "$OUT/htslib-1.9/htslib" ]; then
  echo "[build_ref] unpacking htslib headers"
  tar -xjf "$REF/redist/htslib-1.9.tar.bz2" -C "$OUT" htslib-1.9/htslib
fi

# the image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ statically (exceptions thrown inside a
# dlopen()ed library then crash); use the distribution compiler and the shared libstdc++ instead.
if [ -x /usr/bin/g++ ]; then CXX=/usr/bin/g++; else CXX=g++; fi
# Release flags of the reference: -O3 -fomit-frame-pointer, asserts ON (src/cmake/cxxConfigure.cmake:452-454)
CXXFLAGS="-std=c++11 -O3 -fomit-frame-pointer -fPIC -fvisibility=hidden -ffunction-sections -fdata-sections -w \
  -DBOOST_ERROR_CODE_HEADER_ONLY -DBOOST_SYSTEM_NO_DEPRECATED \
  -I$LIB -I$OUT/boost_1_58_0 -I$OUT/htslib-1.9 -I$HERE/../include -I$HERE"

SRCS=(
  "$LIB/alignment/Alignment.cpp"
  "$LIB/alignment/AlignmentUtil.cpp"
  "$LIB/assembly/IterativeAssembler.cpp"
  "$LIB/assembly/AssembledContig.cpp"
  "$LIB/blt_util/align_path.cpp"
  "$LIB/blt_util/seq_util.cpp"
  "$LIB/blt_util/seq_printer.cpp"
  "$LIB/blt_util/log.cpp"
  "$LIB/blt_util/parse_util.cpp"
  "$LIB/blt_util/blt_exception.cpp"
  "$LIB/blt_util/known_pos_range2.cpp"
  "$LIB/blt_util/qscore_snp.cpp"
  "$LIB/blt_util/qscore_cache.cpp"
  "$LIB/common/Exceptions.cpp"
  "$GSC/SplitReadAlignment.cpp"
  "$LIB/manta/SVScoreInfo.cpp"
  "$OUT/boost_1_58_0/libs/date_time/src/gregorian/greg_month.cpp"
  "$HERE/ref_shim.cpp"
  "$HERE/ref_refiner_shim.cpp"
  # the SV locus graph and the per-bin edge retriever (row a1)
  "$LIB/svgraph/SVLocusSet.cpp"
  "$LIB/svgraph/SVLocus.cpp"
  "$LIB/svgraph/SVLocusNode.cpp"
  "$LIB/svgraph/GenomeInterval.cpp"
  "$LIB/svgraph/EdgeInfo.cpp"
  "$LIB/svgraph/SVLocusSampleCounts.cpp"
  "$GSC/EdgeRetrieverBin.cpp"
  "$LIB/htsapi/bam_header_info.cpp"
  "$HERE/ref_graph_shim.cpp"
  "$HERE/ref_split_shim.cpp"
  "$GSC/SVScorerShared.cpp"
)
# boost.serialization (narrow-character archives only)
for f in "$OUT"/boost_1_58_0/libs/serialization/src/*.cpp; do
  case "$(basename "$f" .cpp)" in
    *w*archive*|*wi*|*wo*|xml_w*|text_w*|basic_xml_grammar|xml_grammar|xml_wgrammar|codecvt_null|utf8_codecvt_facet|shared_ptr_helper) continue ;;
  esac
  SRCS+=("$f")
done

OBJS=()
pids=()
for src in "${SRCS[@]}"; do
  obj="$OUT/obj/$(basename "${src%.cpp}").o"
  OBJS+=("$obj")
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/../include/manta_b200.h" -nt "$obj" ] || [ "$HERE/manta_oracle.h" -nt "$obj" ] ||
     { [ "$(basename "$src")" = ref_graph_shim.cpp ] && [ "$HERE/../include/manta_b200_graph.hpp" -nt "$obj" ]; }; then
    $CXX $CXXFLAGS -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done

# --gc-sections drops getRefAlignment (needs htslib objects we do not build); only ref_* are exported
$CXX -shared -o "$OUT/libmanta_ref.so" "${OBJS[@]}" -Wl,--gc-sections -Wl,--no-undefined -lpthread
echo "[build_ref] built $OUT/libmanta_ref.so"
