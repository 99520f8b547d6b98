"""TEST INFRASTRUCTURE ONLY -- compile the reference's own sources for the hot path into oracle/_ref/:
  libchisel_ref.so  Thirdparty/open_chisel/src/**/*.cpp (TSDF) against oracle/eigen_standin (Eigen is not installed)
  liborb_ref.so     src/ORBextractor.cc against oracle/cv_standin (OpenCV's C++ headers are not installed); its four
                    OpenCV primitives resolve to the C restatements in liboracle.so that are pinned bit-exactly to cv2

The sources are compiled where they lie under /root/reference/Thirdparty/open_chisel (nothing is copied into this
repository); outputs go only to oracle/_ref/ (git-ignored, but it travels to the GPU box with gpurun).  open_chisel's
one external dependency is Eigen, which this image does not have: the sources are compiled against the stand-in
headers in oracle/eigen_standin/ (fixed-size vector/matrix arithmetic restated with Eigen 3.4's evaluation order; see
the header of oracle/eigen_standin/Eigen/Core).  Flags mirror the reference's Ubuntu-24.04 configuration
(config.sh:19-21: no -march=native, hence no FMA contraction).  The reference's own build system (cmake) is not run.

On the GPU box /root/reference does not exist: build() then just returns the prebuilt library (or None).
"""
import pathlib
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference/Thirdparty/open_chisel")
OUTDIR = HERE / "_ref"
OUT = OUTDIR / "libchisel_ref.so"
CXXFLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
            "-I", str(HERE / "eigen_standin"), "-I", str(REF / "include")]


def build(force=False):
    if not REF.exists():
        return str(OUT) if OUT.exists() else None
    srcs = sorted(REF.glob("src/*.cpp")) + sorted(REF.glob("src/*/*.cpp")) + [HERE / "ref_chisel_harness.cpp"]
    deps = srcs + list((HERE / "eigen_standin" / "Eigen").iterdir())
    if OUT.exists() and not force and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(OUT)
    objdir = OUTDIR / "obj"
    objdir.mkdir(parents=True, exist_ok=True)

    def cc(src):
        obj = objdir / (src.stem + ".o")
        subprocess.check_call(["g++"] + CXXFLAGS + ["-c", str(src), "-o", str(obj)])
        return str(obj)

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(cc, srcs))
    subprocess.check_call(["g++", "-shared", "-o", str(OUT)] + objs + ["-lm", "-pthread"])
    shutil.rmtree(objdir, ignore_errors=True)
    return str(OUT)


ORB_OUT = OUTDIR / "liborb_ref.so"


def build_orb(force=False):
    """src/ORBextractor.cc + oracle/ref_orb_harness.cpp -> oracle/_ref/liborb_ref.so (links oracle/liboracle.so)."""
    from . import build as oracle_build
    src = pathlib.Path("/root/reference/src/ORBextractor.cc")
    if not src.exists():
        return str(ORB_OUT) if ORB_OUT.exists() else None
    oracle_so = pathlib.Path(oracle_build.build())
    deps = [src, HERE / "ref_orb_harness.cpp", HERE / "cv_standin" / "opencv2" / "opencv.hpp", oracle_so]
    if ORB_OUT.exists() and not force and all(ORB_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(ORB_OUT)
    OUTDIR.mkdir(parents=True, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
             "-I", str(HERE / "cv_standin"), "-I", "/root/reference/include"]
    subprocess.check_call(["g++"] + flags + ["-shared", "-o", str(ORB_OUT), str(src), str(HERE / "ref_orb_harness.cpp"),
                           "-L", str(HERE), "-l:liboracle.so", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    return str(ORB_OUT)


MATCH_OUT = OUTDIR / "libmatch_ref.so"


def _write_grid_slices():
    """Frame / KeyFrame grid functions -> oracle/_ref/gen/*.inc (build intermediates, git-ignored)"""
    ref = pathlib.Path("/root/reference")
    gen = OUTDIR / "gen"
    gen.mkdir(parents=True, exist_ok=True)
    ftext = (ref / "src" / "Frame.cc").read_text()
    parts = [_slice_function(ftext, "void Frame::AssignFeaturesToGrid()"), _slice_function(ftext, "bool Frame::PosInGrid("),
             _slice_function(ftext, "vector<size_t> Frame::GetFeaturesInArea("), _slice_function(ftext, "void Frame::ComputeStereoFromRGBD(")]
    (gen / "frame_grid_slices.inc").write_text("// generated at build time from /root/reference/src/Frame.cc -- do not commit\n" + "\n\n".join(parts) + "\n")
    ktext = (ref / "src" / "KeyFrame.cc").read_text()
    mtext = (ref / "src" / "MapPoint.cc").read_text()
    (gen / "mappoint_slice.inc").write_text("// generated at build time from /root/reference/src/MapPoint.cc -- do not commit\n" +
                                            _slice_function(mtext, "void MapPoint::ComputeDistinctiveDescriptors()") + "\n")
    (gen / "keyframe_grid_slice.inc").write_text("// generated at build time from /root/reference/src/KeyFrame.cc -- do not commit\n" +
                                                 _slice_function(ktext, "vector<size_t> KeyFrame::GetFeaturesInArea(") + "\n")


def build_match(force=False):
    """src/ORBmatcher.cc + Thirdparty/DBoW2/DBoW2/FeatureVector.cpp + oracle/ref_match_harness.cpp -> oracle/_ref/libmatch_ref.so.
    The PLVS data model (Frame/KeyFrame/MapPoint/GeometricCamera) comes from oracle/plvs_standin/plvs_types.hpp, force-included
    with the include guards of the real headers pre-defined."""
    ref = pathlib.Path("/root/reference")
    src = ref / "src" / "ORBmatcher.cc"
    if not src.exists():
        return str(MATCH_OUT) if MATCH_OUT.exists() else None
    srcs = [src, ref / "Thirdparty" / "DBoW2" / "DBoW2" / "FeatureVector.cpp", HERE / "ref_match_harness.cpp", HERE / "ref_grid_slices.cpp"]
    deps = srcs + [ref / "src" / "Frame.cc", ref / "src" / "KeyFrame.cc", ref / "src" / "MapPoint.cc", HERE / "plvs_standin" / "plvs_types.hpp", HERE / "plvs_standin" / "sophus" / "se3.hpp",
                   HERE / "cv_standin" / "opencv2" / "opencv.hpp", HERE / "eigen_standin" / "Eigen" / "Core", HERE / "eigen_standin" / "Eigen" / "Geometry"]
    if MATCH_OUT.exists() and not force and all(MATCH_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(MATCH_OUT)
    from . import build as oracle_build
    oracle_build.build()
    OUTDIR.mkdir(parents=True, exist_ok=True)
    _write_grid_slices()
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
             "-I", str(HERE / "plvs_standin"), "-I", str(HERE / "cv_standin"), "-I", str(HERE / "eigen_standin"),
             "-I", str(ref / "include"), "-I", str(ref), "-I", str(OUTDIR), "-include", str(HERE / "plvs_standin" / "plvs_types.hpp")]
    subprocess.check_call(["g++"] + flags + ["-shared", "-o", str(MATCH_OUT)] + [str(x) for x in srcs] +
                          ["-L", str(HERE), "-l:liboracle.so", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    shutil.rmtree(OUTDIR / "gen", ignore_errors=True)        # the slices are build intermediates: nothing of the reference's text stays behind
    return str(MATCH_OUT)


STEREO_OUT = OUTDIR / "libstereo_ref.so"


def _slice_function(text, signature):
    """the definition that starts with `signature` up to its matching closing brace (comments / strings in that function hold no braces)"""
    a = text.index(signature)
    i = text.index("{", a)
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[a:j + 1]
    raise RuntimeError("unbalanced braces")


def build_stereo(force=False):
    """Frame::ComputeStereoMatches: sliced out of src/Frame.cc at build time into oracle/_ref/gen/ (git-ignored, never committed) and
    compiled inside oracle/ref_stereo_harness.cpp, together with the reference's ORBextractor.cc and ORBmatcher.cc
    -> oracle/_ref/libstereo_ref.so."""
    ref = pathlib.Path("/root/reference")
    fsrc = ref / "src" / "Frame.cc"
    if not fsrc.exists():
        return str(STEREO_OUT) if STEREO_OUT.exists() else None
    srcs = [fsrc, ref / "src" / "ORBextractor.cc", ref / "src" / "ORBmatcher.cc", ref / "Thirdparty" / "DBoW2" / "DBoW2" / "FeatureVector.cpp",
            HERE / "ref_stereo_harness.cpp", HERE / "ref_grid_slices.cpp"]
    deps = srcs + [ref / "src" / "KeyFrame.cc", HERE / "plvs_standin" / "plvs_types.hpp", HERE / "cv_standin" / "opencv2" / "opencv.hpp"]
    if STEREO_OUT.exists() and not force and all(STEREO_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(STEREO_OUT)
    from . import build as oracle_build
    oracle_build.build()
    gen = OUTDIR / "gen"; obj = OUTDIR / "obj_stereo"
    gen.mkdir(parents=True, exist_ok=True); obj.mkdir(parents=True, exist_ok=True)
    _write_grid_slices()
    body = _slice_function(fsrc.read_text(), "void Frame::ComputeStereoMatches()")
    (gen / "frame_stereo_slice.inc").write_text("// generated at build time from /root/reference/src/Frame.cc -- do not commit\n" + body + "\n")
    base = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w"]
    inc_cv = ["-I", str(HERE / "cv_standin"), "-I", str(ref / "include")]
    inc_all = ["-I", str(HERE / "plvs_standin"), "-I", str(HERE / "cv_standin"), "-I", str(HERE / "eigen_standin"), "-I", str(ref / "include"), "-I", str(ref),
               "-I", str(OUTDIR), "-include", str(HERE / "plvs_standin" / "plvs_types.hpp")]
    objs = []
    for src, inc in ((srcs[1], inc_cv), (srcs[2], inc_all), (srcs[3], inc_all), (srcs[4], inc_all), (srcs[5], inc_all)):
        o = obj / (src.stem + ".o")
        subprocess.check_call(["g++"] + base + inc + ["-c", str(src), "-o", str(o)])
        objs.append(str(o))
    subprocess.check_call(["g++", "-shared", "-o", str(STEREO_OUT)] + objs + ["-L", str(HERE), "-l:liboracle.so", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    shutil.rmtree(OUTDIR / "gen", ignore_errors=True)
    shutil.rmtree(obj, ignore_errors=True)
    return str(STEREO_OUT)


FRUSTUM_OUT = OUTDIR / "libfrustum_ref.so"


def build_frustum(force=False):
    """Frame::isInFrustum + MapPoint::PredictScale / getters + Pinhole::project, sliced out of the reference at build time and compiled
    against oracle/plvs_standin/plvs_frustum_types.hpp -> oracle/_ref/libfrustum_ref.so"""
    ref = pathlib.Path("/root/reference")
    files = [ref / "src" / "Frame.cc", ref / "src" / "MapPoint.cc", ref / "src" / "CameraModels" / "Pinhole.cpp"]
    if not all(f.exists() for f in files):
        return str(FRUSTUM_OUT) if FRUSTUM_OUT.exists() else None
    deps = files + [HERE / "ref_frustum_harness.cpp", HERE / "plvs_standin" / "plvs_frustum_types.hpp", HERE / "eigen_standin" / "Eigen" / "Core"]
    if FRUSTUM_OUT.exists() and not force and all(FRUSTUM_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(FRUSTUM_OUT)
    gen = OUTDIR / "gen"
    gen.mkdir(parents=True, exist_ok=True)
    ftext, mtext, ptext = (f.read_text() for f in files)
    parts = [_slice_function(ftext, "bool Frame::isInFrustum(MapPointPtr& pMP, float viewingCosLimit)"),
             _slice_function(mtext, "int MapPoint::PredictScale(const float &currentDist, Frame* pF)"),
             _slice_function(mtext, "Eigen::Vector3f MapPoint::GetWorldPos()"), _slice_function(mtext, "Eigen::Vector3f MapPoint::GetNormal()"),
             _slice_function(mtext, "float MapPoint::GetMinDistanceInvariance()"), _slice_function(mtext, "float MapPoint::GetMaxDistanceInvariance()"),
             _slice_function(ptext, "Eigen::Vector2f Pinhole::project(const Eigen::Vector3f &v3D) const")]
    (gen / "frustum_slices.inc").write_text("// generated at build time from the reference -- do not commit\n" + "\n\n".join(parts) + "\n")
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", "-I", str(HERE / "plvs_standin"), "-I", str(HERE / "eigen_standin"),
             "-I", str(OUTDIR), "-include", str(HERE / "plvs_standin" / "plvs_frustum_types.hpp")]
    try:
        subprocess.check_call(["g++"] + flags + ["-shared", "-o", str(FRUSTUM_OUT), str(HERE / "ref_frustum_harness.cpp"), "-lm"])
    finally:
        shutil.rmtree(gen, ignore_errors=True)
    return str(FRUSTUM_OUT)


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
    print(build_orb(force="-f" in sys.argv))
    print(build_match(force="-f" in sys.argv))
    print(build_stereo(force="-f" in sys.argv))
    print(build_frustum(force="-f" in sys.argv))


BOW_OUT = OUTDIR / "libbow_ref.so"


def build_bow(force=False):
    """Thirdparty/DBoW2 (TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp, DUtils) + oracle/ref_bow_harness.cpp
    -> oracle/_ref/libbow_ref.so, against the OpenCV stand-in (cv::Mat of bytes; cv::FileStorage only as aborting stubs: vocabularies are
    read with the reference's loadFromTextFile)."""
    ref = pathlib.Path("/root/reference/Thirdparty/DBoW2")
    if not (ref / "DBoW2" / "TemplatedVocabulary.h").exists():
        return str(BOW_OUT) if BOW_OUT.exists() else None
    srcs = [HERE / "ref_bow_harness.cpp"] + [ref / "DBoW2" / n for n in ("BowVector.cpp", "FeatureVector.cpp", "FORB.cpp", "ScoringObject.cpp")] + \
        [ref / "DUtils" / n for n in ("Random.cpp", "Timestamp.cpp")]
    deps = srcs + [ref / "DBoW2" / "TemplatedVocabulary.h", HERE / "cv_standin" / "opencv2" / "opencv.hpp"]
    if BOW_OUT.exists() and not force and all(BOW_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(BOW_OUT)
    OUTDIR.mkdir(parents=True, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", "-I", str(HERE / "cv_standin"), "-I", str(ref / "DBoW2"), "-I", str(ref)]
    subprocess.check_call(["g++"] + flags + ["-shared", "-o", str(BOW_OUT)] + [str(x) for x in srcs] + ["-lm"])
    return str(BOW_OUT)


LINEMATCH_OUT = OUTDIR / "liblinematch_ref.so"


def _slice_class(text, head):
    """the class declaration that starts with `head` up to the `};` that closes it"""
    a = text.index(head)
    i = text.index("{", a)
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[a:text.index(";", j) + 1]
    raise ValueError(head)


def build_linematch(force=False):
    """Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp (multi-index hashing k-NN over 256-bit LBD descriptors, what
    LineMatcher::ComputeDescriptorMatches calls, src/LineMatcher.cc:2572) + oracle/ref_linematch_harness.cpp -> oracle/_ref/liblinematch_ref.so.
    The .cpp is compiled unmodified where it lies.  Its header (descriptor_custom.hpp) also declares the LSD / LBD extractor classes, which need
    half of OpenCV: the build writes a `line_descriptor_custom.hpp` holding only the text of `class BinaryDescriptorMatcher` (sliced out at build
    time into the git-ignored oracle/_ref/gen/, deleted afterwards) and puts it first on the include path."""
    ref = pathlib.Path("/root/reference/Thirdparty/line_descriptor")
    src = ref / "src" / "binary_descriptor_matcher_custom.cpp"
    if not src.exists():
        return str(LINEMATCH_OUT) if LINEMATCH_OUT.exists() else None
    hdr = ref / "include" / "line_descriptor" / "descriptor_custom.hpp"
    deps = [src, hdr, HERE / "ref_linematch_harness.cpp", HERE / "cv_standin" / "opencv2" / "opencv.hpp"]
    if LINEMATCH_OUT.exists() and not force and all(LINEMATCH_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(LINEMATCH_OUT)
    gen = OUTDIR / "gen_linematch"
    gen.mkdir(parents=True, exist_ok=True)
    decl = _slice_class(hdr.read_text(), "class CV_EXPORTS BinaryDescriptorMatcher")
    (gen / "line_descriptor_custom.hpp").write_text(
        "// generated at build time from /root/reference/Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp -- do not commit\n"
        "#pragma once\n#include <map>\n#include <vector>\n#include <cmath>\n#include <cstring>\n#include <opencv2/opencv.hpp>\n"
        "#include \"types_custom.hpp\"\n#ifndef CV_EXPORTS\n#define CV_EXPORTS\n#endif\n"
        "namespace cv { namespace line_descriptor_c {\n" + decl + "\n} }\n")
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", "-I", str(gen), "-I", str(HERE / "cv_standin"), "-I", str(ref / "src")]
    try:
        subprocess.check_call(["g++"] + flags + ["-shared", "-o", str(LINEMATCH_OUT), str(src), str(HERE / "ref_linematch_harness.cpp"), "-lm"])
    finally:
        shutil.rmtree(gen, ignore_errors=True)
    return str(LINEMATCH_OUT)


MAPPLY_OUT = OUTDIR / "libmapply_ref.so"


def build_mapply(force=False):
    """PointCloudMap<PointT>::WritePLY + writeCustomData (src/PointCloudMap.cc:304-437: the volumetric map file PointCloudMapChisel::SaveMap writes and
    LoadMap reads back), sliced out at build time into the git-ignored oracle/_ref/gen_mapply/ and compiled inside oracle/ref_mapply_harness.cpp
    -> oracle/_ref/libmapply_ref.so."""
    src = pathlib.Path("/root/reference/src/PointCloudMap.cc")
    if not src.exists():
        return str(MAPPLY_OUT) if MAPPLY_OUT.exists() else None
    deps = [src, HERE / "ref_mapply_harness.cpp"]
    if MAPPLY_OUT.exists() and not force and all(MAPPLY_OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(MAPPLY_OUT)
    text = src.read_text()
    gen = OUTDIR / "gen_mapply"
    gen.mkdir(parents=True, exist_ok=True)
    parts = [_slice_function(text, "template <class PointT, typename std::enable_if<!pcl::traits::has_field<PointT, pcl::fields::kfid>::value>::type* = nullptr>"),
             _slice_function(text, "template <class PointT, typename std::enable_if<pcl::traits::has_field<PointT, pcl::fields::kfid>::value>::type* = nullptr>"),
             "template<typename PointT>\n" + _slice_function(text, "bool PointCloudMap<PointT>::WritePLY(")]
    (gen / "mapply_slices.inc").write_text("// generated at build time from /root/reference/src/PointCloudMap.cc -- do not commit\n" + "\n\n".join(parts) + "\n")
    try:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", "-I", str(gen), "-shared", "-o", str(MAPPLY_OUT),
                               str(HERE / "ref_mapply_harness.cpp")])
    finally:
        shutil.rmtree(gen, ignore_errors=True)
    return str(MAPPLY_OUT)
