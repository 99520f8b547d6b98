/* plvs_b200.h -- C ABI of libplvs_b200.so: the B200-native (sm_100a) replacement for the
 * per-frame data-parallel hot path of PLVS (luigifreda/plvs):
 *
 *   ORB extraction      PLVS2::ORBextractor::operator()            src/ORBextractor.cc:1245
 *   Hamming matching    PLVS2::ORBmatcher::SearchByProjection x2   src/ORBmatcher.cc:71, :1774
 *                       PLVS2::ORBmatcher::SearchForTriangulation  src/ORBmatcher.cc:999
 *                       PLVS2::ORBmatcher::DescriptorDistance      src/ORBmatcher.cc:2198
 *   TSDF integration    chisel::Chisel::IntegrateDepthScan[ColorWithOneCameraModelBGR]
 *                       Thirdparty/open_chisel/include/open_chisel/Chisel.h:68, :198
 *                       (reached through chisel_server::ChiselServer::IntegrateLastDepthImage,
 *                        Thirdparty/chisel_server/src/ChiselServer.cpp:632)
 *
 * The reference has no FFI layer: its boundary is the three C++ class surfaces above.  The
 * header-only shims in shim/ keep those signatures and forward to this ABI.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (PLVS_OK) or a negative
 * PLVS_E* code and never aborts; all buffers are caller-owned unless stated; every handle owns a
 * CUDA stream + workspace and calls on DISTINCT handles are thread-safe (the reference runs
 * left/right extractors and Tracking/LocalMapping matchers on concurrent threads).  There is no
 * CPU fallback: without a CUDA device every create() fails with PLVS_ENODEV.
 */
#ifndef PLVS_B200_H_
#define PLVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLVS_OK 0
#define PLVS_EINVAL (-1)   /* bad argument */
#define PLVS_ENODEV (-2)   /* no CUDA device / CUDA runtime failure */
#define PLVS_ENOMEM (-3)   /* allocation failed */
#define PLVS_ECAP (-4)     /* caller capacity too small (n_out holds the required size) */
#define PLVS_ESTATE (-5)   /* missing camera/pose (mirrors ChiselServer gotInfo/gotPose guards) */

#define PLVS_MAX_LEVELS 16

const char* plvs_version(void);
const char* plvs_last_error(void);          /* thread-local description of the last failure */
int plvs_device_count(void);
/* Per-kernel CUDA-event timing on each handle's stream (off by default; used by bench.py for the roofline
 * figures).  plvs_*_kernel_times() return accumulated milliseconds and launch counts since the last reset. */
/* mask: 1 = ORB kernels, 2 = matcher kernels, 4 = TSDF kernels, 8 = restrict to the roofline kernel (k_integrate); 0 = off */
int plvs_set_profiling(int mask);
#define PLVS_ORB_K_RESIZE 0
#define PLVS_ORB_K_FAST 1
#define PLVS_ORB_K_COMPACT 2
#define PLVS_ORB_K_BLUR 3
#define PLVS_ORB_K_DESCRIBE 4
#define PLVS_ORB_K_DISTRIBUTE 5
#define PLVS_ORB_K_GRID 6
#define PLVS_MATCH_K_GRID 0
#define PLVS_MATCH_K_CANDIDATES 1
#define PLVS_MATCH_K_RESOLVE 2
#define PLVS_MATCH_K_TRIANGULATE 3
#define PLVS_MATCH_K_FUSE 4
#define PLVS_MATCH_K_BOW 5
#define PLVS_MATCH_K_INIT 6
#define PLVS_MATCH_K_LINES 7
#define PLVS_TSDF_K_TILES 0
#define PLVS_TSDF_K_CLASSIFY 1
#define PLVS_TSDF_K_INTEGRATE 2
#define PLVS_TSDF_K_COMMIT 3
#define PLVS_TSDF_K_MESH 4
#define PLVS_TSDF_K_BIND 5
#define PLVS_K_SLOTS 12
/* bytes the library has moved over the bus since the last reset (process-wide): host->device and device->host, counted at every copy the
 * library issues and at every result a kernel writes straight into mapped host memory */
int plvs_io_bytes(long long* h2d, long long* d2h, int reset);
/* Kernels of `device` may read the memory of `peer` over NVLink (cudaDeviceEnablePeerAccess): the two-GPU eye split of a stereo frame
 * (src/Frame.cc:314-329 extracts the two eyes in two threads) keeps the right eye's pyramid and keypoints on the second GPU and
 * lets plvs_stereo_match on the first one read them in place. */
int plvs_enable_peer_access(int device, int peer);
/* pinned host memory for the e2e path (cudaHostAlloc / cudaFreeHost) */
int plvs_host_alloc(void** p, size_t bytes);
int plvs_host_free(void* p);

/* ------------------------------------------------------------------------------------------ */
/* ORB extraction -- replaces PLVS2::ORBextractor (include/ORBextractor.h:59-170)              */
/* ------------------------------------------------------------------------------------------ */
typedef struct plvs_orb plvs_orb;

/* ctor arguments of ORBextractor (include/ORBextractor.h:78) */
typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
} plvs_orb_params;

/* binary layout of cv::KeyPoint (28 bytes): the shim passes std::vector<cv::KeyPoint>::data() */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} plvs_keypoint;

int plvs_orb_create(const plvs_orb_params* p, int device, plvs_orb** out);
void plvs_orb_destroy(plvs_orb* h);

/* ORBextractor::operator() (src/ORBextractor.cc:1245-1389).  gray: CV_8UC1, `stride` bytes per row;
 * on_device!=0 => gray is a device pointer.  (lap0,lap1) = vLappingArea.  kps/desc have room for
 * `cap` keypoints (desc is cap x 32 bytes).  *n_out = number of keypoints, *mono_index_out = the
 * value operator() returns.  Returns PLVS_ECAP (with *n_out = needed) if cap is too small. */
int plvs_orb_extract(plvs_orb* h, const uint8_t* gray, int w, int h_, int stride, int on_device,
                     int lap0, int lap1, plvs_keypoint* kps, uint8_t* desc, int cap,
                     int* n_out, int* mono_index_out);

/* Same, for `batch` independent frames of equal size in one pass (frame b starts at
 * gray + b*frame_stride; outputs at kps + b*cap, desc + b*cap*32, n_out[b], mono_index_out[b]). */
int plvs_orb_extract_batch(plvs_orb* h, int batch, const uint8_t* gray, int w, int h_, int stride,
                           size_t frame_stride, int on_device, int lap0, int lap1,
                           plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index_out);

/* getters of ORBextractor (include/ORBextractor.h:90-113); arrays hold nlevels entries */
int plvs_orb_tables(const plvs_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int32_t* features_per_level);

/* mvImagePyramid / mvImagePyramidFiltered (include/ORBextractor.h:125-127) of frame `frame` of the
 * last batch: device pointer + geometry, and a host download for callers that read the pyramid
 * (Frame::ComputeStereoMatches, src/Frame.cc:1789). */
int plvs_orb_pyramid_level(const plvs_orb* h, int frame, int level, int blurred,
                           const uint8_t** dptr, int* w, int* h_, int* pitch);
int plvs_orb_download_level(plvs_orb* h, int frame, int level, int blurred, uint8_t* host, int host_stride);

/* Device-resident result of frame `frame` of the last batch (valid until the next extract on this
 * handle; only when no keypoint fell in the lapping area): lets the matcher run without a
 * device->host->device round trip. */
/* Step before extraction: the same extraction on 3- or 4-channel 8-bit frames; cv::cvtColor(..,
 * COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) (src/Tracking.cc:1797-1810, OpenCV 4 fixed point) runs on the device in front of the pyramid.
 * stride / frame_stride in bytes; is_rgb = Tracking::mbRGB. */
int plvs_orb_extract_batch_color(plvs_orb* h, int batch, const uint8_t* img, int w, int h_, int stride, size_t frame_stride, int nch, int is_rgb,
                                 int on_device, int lap0, int lap1, plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index_out);

/* Step after extraction (§8f rank 2): Frame::ComputeStereoFromRGBD (src/Frame.cc:2251-2279) for frame `frame` of the last batch,
 * on the device-resident keypoints.  depth: CV_32F image in metres (host, or device when on_device), row stride in bytes.
 * keys_un_x: mvKeysUn[i].pt.x (host, n floats) or NULL when the camera has no distortion (mvKeysUn == mvKeys).  uright / depth_out
 * (host, n floats, nullable) receive mvuRight / mvDepth; *d_uright (nullable) the device copy of mvuRight, valid until the next
 * extraction on this handle: pass it as plvs_frame_view.uright with on_device = PLVS_VIEW_ON_DEVICE. */
int plvs_orb_stereo_from_rgbd(plvs_orb* h, int frame, const float* depth, int w, int h_, int stride_bytes, int on_device, float bf,
                              const float* keys_un_x, float* uright, float* depth_out, const float** d_uright);

/* Step after extraction (§8f rank 2): Frame::UndistortKeyPoints (src/Frame.cc:1507-1553) on the device-resident keypoints of frame `frame`:
 * cv::undistortPoints(pts, K, distCoef, Mat(), K) (five iterations in double, pinhole model).  K = fx, fy, cx, cy; dist = the ndist <= 14 float
 * coefficients of mDistCoef (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1..s4]]]); dist[0] == 0 copies the keypoints like the reference.  keys_un (host, n
 * records, nullable) receives mvKeysUn; *d_keys_un (nullable) the device copy, valid until the next extraction: use it as plvs_frame_view.keys. */
int plvs_orb_undistort(plvs_orb* h, int frame, const float K[4], const float* dist, int ndist, plvs_keypoint* keys_un, const plvs_keypoint** d_keys_un);

typedef struct {
    int32_t n;
    const plvs_keypoint* keys;   /* device */
    const uint8_t* desc;         /* device, n x 32 */
    uint64_t cache_key;          /* unique per (handle, extract call, frame): pass it on in plvs_frame_view.cache_key */
    const int32_t* grid_cell_start;   /* device; the frame's feature grid when plvs_orb_set_frame_grid is in effect, else NULL: */
    const int32_t* grid_sorted;       /* pass both on in plvs_frame_view.grid_cell_start / grid_sorted */
} plvs_orb_device_view;

/* Frame::AssignFeaturesToGrid (src/Frame.cc:716-746) as the last step of frame construction, where the reference has it (Frame constructor,
 * src/Frame.cc:598): with bounds set, every extraction also bins the device-resident keypoints of each frame into the 64 x 48 grid
 * (mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv: the static members of Frame, src/Frame.cc:444-449,1749-1778)
 * and plvs_orb_device_result hands the grid out.  Valid when mvKeysUn == mvKeys (no distortion: RGB-D / rectified input); a matcher given
 * the grid skips its own build.  bounds == NULL turns it off again. */
int plvs_orb_set_frame_grid(plvs_orb* h, const float bounds[6]);
int plvs_orb_device_result(const plvs_orb* h, int frame, plvs_orb_device_view* out);

/* Inspection: FAST candidates of (frame, level) of the last batch in the order they enter
 * DistributeOctTree (vToDistributeKeys, src/ORBextractor.cc:867-998): x, y in level pixel
 * coordinates (border included) and the FAST score.  Used by the stage-parity tests. */
int plvs_orb_candidates(const plvs_orb* h, int frame, int level, int32_t* x, int32_t* y, int32_t* score,
                        int cap, int* n_out);

/* statistics of the last batch (for bench/roofline accounting) */
typedef struct {
    int64_t pyramid_pixels;      /* P of SURVEY.md §8: sum over levels of w*h, one frame */
    int64_t candidates;          /* FAST candidates after per-cell NMS, summed over the batch */
    int64_t keypoints;           /* keypoints, summed over the batch */
    int32_t kernel_launches;     /* kernels launched by the last extract call */
    float host_wait_candidates_ms;   /* host blocked until FAST candidates arrived (copy-in + pyramid + FAST + compaction) */
    float host_distribute_ms;        /* DistributeOctTree on host threads, whole batch */
    float host_wait_describe_ms;     /* host blocked on blur (overlapped) + orientation/descriptor kernel */
    float host_assemble_ms;          /* copy-out into the caller's vectors */
} plvs_orb_stats;
int plvs_orb_last_stats(const plvs_orb* h, plvs_orb_stats* out);
int plvs_orb_kernel_times(plvs_orb* h, float* ms /*PLVS_K_SLOTS*/, int32_t* launches /*PLVS_K_SLOTS*/, int reset);

/* ------------------------------------------------------------------------------------------ */
/* Matching -- replaces the three ORBmatcher entry points + DescriptorDistance                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct plvs_match plvs_match;

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2198-2225): host helper, 32-byte descriptors */
int plvs_hamming256(const uint8_t* a, const uint8_t* b);

/* Flat view of the Frame / KeyFrame members the matchers read (RGB-D / rectified stereo: Nleft==-1).
 * Pointers are host pointers unless on_device != 0: on_device = PLVS_VIEW_ON_DEVICE means keys / desc / uright are device
 * pointers (e.g. plvs_orb_device_result); or-ing PLVS_VIEW_URIGHT_ON_HOST says that `uright` alone is still a host array
 * (mvuRight as Frame::ComputeStereoFromRGBD leaves it on the CPU), which is then staged per call. */
#define PLVS_VIEW_ON_DEVICE 1
#define PLVS_VIEW_URIGHT_ON_HOST 2
typedef struct {
    int32_t n;                    /* Frame::N */
    const plvs_keypoint* keys;    /* Frame::mvKeysUn */
    const uint8_t* desc;          /* Frame::mDescriptors, n x 32 */
    const float* uright;          /* Frame::mvuRight (NULL => all -1) */
    float min_x, min_y, max_x, max_y;     /* mnMinX.. (src/Frame.cc:1749-1778) */
    float grid_inv_w, grid_inv_h;         /* mfGridElementWidthInv / HeightInv (src/Frame.cc:448-449) */
    float scale_factors[PLVS_MAX_LEVELS]; /* Frame::mvScaleFactors */
    float level_sigma2[PLVS_MAX_LEVELS];  /* KeyFrame::mvLevelSigma2 (triangulation only) */
    int32_t nlevels;
    float bf;                             /* Frame::mbf */
    int32_t on_device;
    uint64_t cache_key;                   /* 0 = none.  Same non-zero key on consecutive searches of one handle = same keypoints:
                                             the feature grid built for the previous search is reused (the reference builds it once
                                             per Frame, src/Frame.cc:598) */
    const int32_t* grid_cell_start;       /* optional, device (with on_device): the frame's grid as plvs_orb_device_result returned it */
    const int32_t* grid_sorted;           /* (built at frame construction); both NULL = the matcher builds the grid itself */
} plvs_frame_view;

#define PLVS_Q_OBS_POSITIVE 1u   /* MapPoint::Observations() > 0 */

/* one in-view local map point, as Frame::isInFrustum leaves it (src/Frame.cc:1006-1014) */
typedef struct {
    float proj_x, proj_y, proj_xr;   /* mTrackProjX / Y / XR */
    float track_depth;               /* mTrackDepth */
    float view_cos;                  /* mTrackViewCos */
    int32_t level;                   /* mnTrackScaleLevel */
    uint32_t flags;                  /* PLVS_Q_* */
    uint8_t desc[32];                /* MapPoint::GetDescriptor() */
} plvs_mp_query;

/* one last-frame feature with a non-outlier MapPoint, already projected by the caller with the
 * reference's own expressions (src/ORBmatcher.cc:1804-1827): uv = project(Tcw*Xw), invz = 1/zc */
typedef struct {
    float u, v, invz;
    int32_t last_octave;
    float angle;                     /* last-frame keypoint angle (rotation histogram) */
    uint32_t flags;                  /* PLVS_Q_* */
    uint8_t desc[32];
} plvs_last_query;

int plvs_match_create(int device, plvs_match** out);
void plvs_match_destroy(plvs_match* h);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPointPtr>&, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:71-244), left/RGB-D branch.  Queries are the map points with mbTrackInView,
 * in mvpLocalMapPoints order.  claimed_in[i]!=0 <=> F.mvpMapPoints[i] && Observations()>0 at call
 * time (NULL => none).  assign[i] (host, n entries) receives the index of the query written into
 * F.mvpMapPoints[i] by this call, or -1.  *nmatches = the function's return value. */
int plvs_match_projection_map(plvs_match* h, const plvs_frame_view* F, const plvs_mp_query* q, int nq,
                              float th, float nn_ratio, int far_points, float th_far,
                              const uint8_t* claimed_in, int32_t* assign, int* nmatches);

/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (src/ORBmatcher.cc:1774-1993).
 * forward/backward = bForward/bBackward (:1793-1794) computed by the caller from tlc and mb.
 * assign[i] = query index finally held by CurrentFrame.mvpMapPoints[i] through this call, or -1. */
int plvs_match_projection_last(plvs_match* h, const plvs_frame_view* cur, const plvs_last_query* q, int nq,
                               float th, int forward, int backward, int check_orientation,
                               const uint8_t* claimed_in, int32_t* assign, int* nmatches);

int plvs_match_kernel_times(plvs_match* h, float* ms, int32_t* launches, int reset);
/* rounds of the claim fixed point and kernels launched by the last projection search */
int plvs_match_last_stats(const plvs_match* h, int* rounds, int* kernel_launches);
/* candidate-list re-evaluations of the last projection search after round 0 (-1: the cluster kernel ran, which re-walks every list every round) */
int plvs_match_last_walks(const plvs_match* h);
/* inspection: SM cycles thread 0 of the last one-CTA claim resolution spent in [0] building the claim tables, [1] comparing watch sets, [2] re-evaluating
 * queries, and [3] from the first round to the end of the kernel (rounds + wrap-up) */
int plvs_match_last_phase_cycles(const plvs_match* h, int32_t out[4]);

/* DBoW2::FeatureVector flattened: sorted node ids, CSR offsets, feature indices (ascending per node) */
typedef struct {
    int32_t n_nodes;
    const uint32_t* node_ids;     /* n_nodes, strictly increasing */
    const int32_t* offsets;       /* n_nodes + 1 */
    const int32_t* features;      /* offsets[n_nodes] */
} plvs_featvec;

/* ---------------------------------------------------------------------------------------------------------------------------------
 * Bag of words: ORBVocabulary (include/ORBVocabulary.h = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) as
 * Frame::ComputeBoW / KeyFrame::ComputeBoW use it (src/Frame.cc:1498-1505): transform(vCurrentDesc, mBowVec, mFeatVec, 4).
 * plvs_voc_load_text = TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1436, the ORBvoc.txt format;
 * empty lines are ignored).  plvs_voc_create takes the same tree as flat arrays: node 0 is the root, parent[i] < i, word_id[i] >= 0 for
 * leaves (-1 otherwise), 32 descriptor bytes and a weight per node; scoring / weighting are DBoW2's enums (BowVector.h:39-56).
 * plvs_voc_transform = transform(features, BowVector&, FeatureVector&, levelsup) (:1140-1207): per feature the word, its weight and the node
 * `levelsup` levels above the leaves (any of the three may be NULL); the BowVector as ascending word ids + values (n entries at most); the
 * FeatureVector flattened as in plvs_featvec -- fv_nodes and fv_features hold n entries at most, fv_offsets n + 1.  fv_device, if not NULL, receives the
 * same FeatureVector as device pointers -- valid until the next transform on this handle -- for the matcher entry points that take
 * device-resident frames.  desc may be a device pointer (desc_on_device != 0: the extractor's descriptor buffer). */
typedef struct plvs_voc plvs_voc;
int plvs_voc_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const int32_t* word_id,
                    const uint8_t* desc, const double* weight, plvs_voc** out);
int plvs_voc_load_text(const char* path, int device, plvs_voc** out);
void plvs_voc_destroy(plvs_voc* h);
int plvs_voc_size(const plvs_voc* h);
int plvs_voc_transform(plvs_voc* h, const uint8_t* desc, int n, int desc_on_device, int levelsup, uint32_t* word, double* weight, uint32_t* node,
                       uint32_t* bow_ids, double* bow_vals, int* n_bow, uint32_t* fv_nodes, int32_t* fv_offsets, int32_t* fv_features,
                       int* n_fv_nodes, plvs_featvec* fv_device);
/* the BowVector step of the transform on its own (host arithmetic, no device involved): bow_ids / bow_vals sized for n entries */
int plvs_bow_vector(int scoring, int weighting, const uint32_t* word, const double* weight, int n, uint32_t* bow_ids, double* bow_vals, int* n_bow);

/* ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:999-1242), monocular-camera branch
 * (mpCamera2 == NULL).  has_mp*[i]!=0 <=> KeyFrame::GetMapPoint(i) != NULL.  F12 = the matrix
 * Pinhole::epipolarConstrain builds (src/CameraModels/Pinhole.cpp:125-131), row-major; ep = epipole
 * in KF2 (:1010-1011).  match12[i] (n1 entries) = matched index in KF2 or -1 (== vMatches12). */
int plvs_match_triangulation(plvs_match* h, const plvs_frame_view* kf1, const plvs_frame_view* kf2,
                             const plvs_featvec* fv1, const plvs_featvec* fv2,
                             const uint8_t* has_mp1, const uint8_t* has_mp2,
                             const float F12[9], const float ep[2],
                             int only_stereo, int coarse, int check_orientation,
                             int32_t* match12, int* nmatches);

/* Frame::isInFrustum(MapPointPtr&, viewingCosLimit) (src/Frame.cc:955-1017, RGB-D / rectified stereo) for a whole local map at once --
 * the loop of Tracking::SearchLocalPoints that produces the queries of plvs_match_projection_map.
 * plvs_map_point: GetWorldPos(), GetNormal(), mfMinDistance, mfMaxDistance (the 0.8 / 1.2 factors of Get{Min,Max}DistanceInvariance are
 * applied inside, src/MapPoint.cc:569-579), flags = PLVS_Q_*, desc = GetDescriptor().  plvs_frustum: mRcw (row-major), mtcw, mOw,
 * Pinhole parameters, mbf, the viewing-cosine limit, mfScaleFactor / mnScaleLevels (PredictScale, src/MapPoint.cc:598-613; evaluated
 * without a device logarithm: level thresholds are computed on the host with the host's logf), image bounds mnMinX..mnMaxY.
 * queries[i] receives mTrackProjX/Y/XR, mTrackDepth, mTrackViewCos, mnTrackScaleLevel of point i; in_view[i] = mbTrackInView; the
 * caller hands the in-view entries to plvs_match_projection_map in order.  *n_in_view = number of points in view. */
typedef struct {
    float xw[3], normal[3];
    float min_dist, max_dist;
    uint32_t flags;
    uint8_t desc[32];
} plvs_map_point;
typedef struct {
    float Rcw[9], tcw[3], Ow[3];
    float fx, fy, cx, cy;
    float bf, viewing_cos_limit, scale_factor;
    int32_t nlevels;
    float min_x, min_y, max_x, max_y;
} plvs_frustum;
int plvs_match_in_frustum(plvs_match* h, const plvs_frustum* fr, const plvs_map_point* pts, int n,
                          plvs_mp_query* queries, uint8_t* in_view, int* n_in_view);
/* ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, bFarPoints, thFarPoints) on the queries the last plvs_match_in_frustum call of this
 * handle left on the device (the in-view points, in order): no host round trip between the two steps of Tracking::SearchLocalPoints.
 * assign[i] = index into the map-point array given to plvs_match_in_frustum, or -1.  Otherwise as plvs_match_projection_map. */
int plvs_match_projection_map_resident(plvs_match* h, const plvs_frame_view* F, float th, float nn_ratio, int far_points, float th_far,
                                       const uint8_t* claimed_in, int32_t* assign, int* nmatches);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFramePtr& pKF, const set<MapPointPtr>& sAlreadyFound, th, ORBdist)
 * (src/ORBmatcher.cc:1996-2122), Tracking::Relocalization.  One query per keyframe map point that is not bad, not in
 * sAlreadyFound and passed the caller-side projection / distance gates (:2024-2046): u, v = project(Tcw * Xw), invz unused (set
 * it >= 0), last_octave = PredictScale(dist3D, &CurrentFrame), angle = pKF->mvKeysUn[i].angle, flags = PLVS_Q_OBS_POSITIVE.
 * claimed_in[i] != 0 <=> CurrentFrame.mvpMapPoints[i] is non-null (ANY map point blocks here, :2066).  assign / nmatches as in
 * plvs_match_projection_last. */
int plvs_match_projection_reloc(plvs_match* h, const plvs_frame_view* cur, const plvs_last_query* q, int nq, float th, int orb_dist,
                                int check_orientation, const uint8_t* claimed_in, int32_t* assign, int* nmatches);

/* ORBmatcher::SearchByProjection(KeyFramePtr& pKF, Sophus::Sim3f& Scw, const vector<MapPointPtr>& vpPoints, vector<MapPointPtr>&
 * vpMatched, int th, float ratioHamming) (src/ORBmatcher.cc:509-615) and its sibling with vpPointsKFs / vpMatchedKF (:617-730),
 * LoopClosing.  One query per candidate map point that passed the caller-side gates (:531-566): u, v, last_octave = predicted
 * level, flags = PLVS_Q_OBS_POSITIVE, invz >= 0, angle unused.  matched_in[i] != 0 <=> vpMatched[i] is non-null on entry.
 * assign[i] = query written into vpMatched[i] during the call or -1; *nmatches = the return value. */
int plvs_match_projection_sim3(plvs_match* h, const plvs_frame_view* kf, const plvs_last_query* q, int nq, float th, float ratio_hamming,
                               const uint8_t* matched_in, int32_t* assign, int* nmatches);

/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize)
 * (src/ORBmatcher.cc:732-852; Tracking::MonocularInitialization, src/Tracking.cc:3003-3004, matcher (0.9, true), window 100).
 * prev_matched: f1->n cv::Point2f (x, y), in/out -- the positions of the final matches are written back like the reference does;
 * matches12[i1] (f1->n entries) = index in F2 or -1; *nmatches = the return value. */
int plvs_match_initialization(plvs_match* h, const plvs_frame_view* f1, const plvs_frame_view* f2, float* prev_matched, int window_size,
                              float nn_ratio, int check_orientation, int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFramePtr& pKF, Frame& F, vector<MapPointPtr>& vpMapPointMatches) (src/ORBmatcher.cc:300-506),
 * RGB-D / rectified stereo (Nleft == -1): Tracking::TrackReferenceKeyFrame and Relocalization.  fv_kf / fv_f = pKF->mFeatVec /
 * F.mFeatVec flattened; has_mp_kf[i] != 0 <=> pKF->GetMapPointMatches()[i] is non-null and not bad; nn_ratio / check_orientation
 * = the ORBmatcher constructor arguments.  match_f[i] (f->n entries) = keyframe feature whose map point was assigned to frame
 * feature i (== vpMapPointMatches[i] = vpMapPointsKF[match_f[i]]) or -1; *nmatches = the value SearchByBoW returns. */
int plvs_match_bow(plvs_match* h, const plvs_frame_view* kf, const plvs_frame_view* f,
                   const plvs_featvec* fv_kf, const plvs_featvec* fv_f, const uint8_t* has_mp_kf,
                   float nn_ratio, int check_orientation, int32_t* match_f, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFramePtr& pKF1, KeyFramePtr& pKF2, vector<MapPointPtr>& vpMatches12) (src/ORBmatcher.cc:853-997,
 * LoopClosing / merging): has_mp*[i] != 0 <=> the keyframe's map point i is non-null and not bad.  match12[i1] (kf1->n entries) =
 * index in KF2 whose map point is assigned to feature i1 (vpMatches12[i1] = vpMapPoints2[match12[i1]]) or -1. */
int plvs_match_bow_kf(plvs_match* h, const plvs_frame_view* kf1, const plvs_frame_view* kf2,
                      const plvs_featvec* fv1, const plvs_featvec* fv2, const uint8_t* has_mp1, const uint8_t* has_mp2,
                      float nn_ratio, int check_orientation, int32_t* match12, int* nmatches);

/* ORBmatcher::Fuse(KeyFramePtr& pKF, const vector<MapPointPtr>&, th, bRight=false) (src/ORBmatcher.cc:1244-1435), the
 * search part (:1340-1406).  One query = one map point that passed the caller-side gates (:1277-1338: not bad, not already in
 * the keyframe, positive depth, inside the image, distance range, viewing angle), projected with the reference's own
 * expressions: (u, v) = pCamera->project(Tcw * Xw), ur = u - bf / z, level = pMP->PredictScale(dist3D, pKF), desc =
 * pMP->GetDescriptor().  The search does not read the keyframe's map points, so queries are independent; per query it
 * returns the keypoint the reference selects -- best_idx[i] (-1 if no candidate passed the level / chi-square gates) and
 * best_dist[i] (256 then).  The caller applies `bestDist <= TH_LOW` and the Replace / AddObservation / AddMapPoint
 * bookkeeping in query order (:1409-1427; pointer graph, stays on the host).  *nfused = number of queries with
 * best_dist <= 50 (== the value Fuse returns).  inv_level_sigma2 = KeyFrame::mvInvLevelSigma2 (nlevels entries). */
typedef struct {
    float u, v, ur;
    int32_t level;                   /* nPredictedLevel */
    uint8_t desc[32];
} plvs_fuse_query;
int plvs_match_fuse(plvs_match* h, const plvs_frame_view* kf, const float* inv_level_sigma2,
                    const plvs_fuse_query* q, int nq, float th,
                    int32_t* best_idx, int32_t* best_dist, int* nfused);
/* ORBmatcher::Fuse(KeyFramePtr& pKF, Sophus::Sim3f& Scw, const vector<MapPointPtr>& vpPoints, float th, vector<MapPointPtr>&
 * vpReplacePoint) (src/ORBmatcher.cc:1437-1553, LoopClosing::SearchAndFuse): the same search without the chi-square gate
 * (:1514-1531); `ur` of the queries is ignored.  The caller applies bestDist <= TH_LOW and fills vpReplacePoint / adds the
 * observation (:1534-1548). */
int plvs_match_fuse_sim3(plvs_match* h, const plvs_frame_view* kf, const plvs_fuse_query* q, int nq, float th,
                         int32_t* best_idx, int32_t* best_dist, int* nfused);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:395-461), batched (LocalMapping calls it for every new / fused map
 * point): desc = the observed descriptors of all points back to back (host, 32 bytes each, in the order the caller collected
 * them from mObservations), offsets[n_points+1] = start of each point's run.  best[i] = index INSIDE the run of the descriptor
 * with the least median distance to the others (median = sorted[0.5*(N-1)], first minimum wins), -1 for an empty run. */
int plvs_distinctive_descriptors(plvs_match* h, const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best);

/* SURVEY.md §8f rank 4, line features, the matcher's data-parallel part: LineMatcher::ComputeDescriptorMatches (src/LineMatcher.cc:2567-2615) =
 * cv::line_descriptor_c::BinaryDescriptorMatcher::knnMatch(query, train, lmatches, 2, queryMask, true)
 * (Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp:258-337) + the ratio test `d0 < nn_ratio * d1`.
 * query / train: nq / nt 256-bit LBD descriptors (32 bytes each, host); mask: nq bytes or NULL (0 = the query is skipped; compact result: it gets no
 * row).  Per row r (the queries with a non-zero mask entry, in order): query_idx[r], train_idx[2r], train_idx[2r+1] (nearest, second nearest -- among
 * equal Hamming distances in the order the library's multi-index hashing returns them), dist[2r], dist[2r+1] (Hamming distances as float, like
 * DMatch::distance), valid[r] = vValidMatch.  Output arrays hold nq rows.  Needs nt >= 2 (with fewer train descriptors than k the reference returns
 * uninitialised indices). */
int plvs_line_knn2(plvs_match* h, const uint8_t* query, int nq, const uint8_t* train, int nt, const uint8_t* mask, float nn_ratio,
                   int32_t* query_idx, int32_t* train_idx, float* dist, uint8_t* valid, int* n_rows, int* n_valid);

/* Device view of one frame's pyramid (all levels) of an extractor handle: what Frame::ComputeStereoMatches
 * reads through mpORBextractorLeft/Right->mvImagePyramid (src/Frame.cc:1886,1914). */
typedef struct {
    const uint8_t* data[PLVS_MAX_LEVELS];    /* device pointers */
    int32_t w[PLVS_MAX_LEVELS], h[PLVS_MAX_LEVELS], pitch[PLVS_MAX_LEVELS];
    int32_t nlevels;
} plvs_pyramid_view;
int plvs_orb_pyramid_view(const plvs_orb* h, int frame, int blurred, plvs_pyramid_view* out);

/* Frame::ComputeStereoMatches (src/Frame.cc:1780-1983), rectified stereo: row-band Hamming search (best < 75),
 * 11x11 SAD refinement over +-5 px on the unblurred level of the left keypoint, parabola sub-pixel fit, depth =
 * bf/disparity, median-based outlier cut.  left/right: views whose `keys` are mvKeys / mvKeysRight (distorted
 * coordinates are what the reference uses here) with scale_factors filled; inv_scale = mvInvScaleFactors.
 * uright/depth (host, left->n entries) receive mvuRight / mvDepth (-1 where unmatched). */
int plvs_stereo_match(plvs_match* h, const plvs_frame_view* left, const plvs_frame_view* right,
                      const plvs_pyramid_view* pyr_left, const plvs_pyramid_view* pyr_right, const float* inv_scale,
                      float mb, float mbf, float* uright, float* depth, int* n_valid);

/* ------------------------------------------------------------------------------------------ */
/* TSDF -- replaces chisel_server::ChiselServer's integrate path (ChiselServer.h:81-322)       */
/* ------------------------------------------------------------------------------------------ */
typedef struct plvs_tsdf plvs_tsdf;

/* chisel_server::ChiselServerParams (Thirdparty/chisel_server/src/ChiselServer.cpp:44-69) as
 * PointCloudMapChisel fills it (src/PointCloudMapChisel.cc:46-61); chunk size is 16^3. */
typedef struct {
    float voxel_resolution;
    float trunc_quad, trunc_linear, trunc_const, trunc_scale;
    float weight;
    int32_t use_carving;
    float carving_dist;
    int32_t use_color;
    float near_plane, far_plane;
    int32_t max_blocks;          /* capacity of the device block pool (48 KiB per block) */
} plvs_tsdf_params;

void plvs_tsdf_default_params(plvs_tsdf_params* p);   /* the defaults of ChiselServerParams() */
int plvs_tsdf_create(const plvs_tsdf_params* p, int device, plvs_tsdf** out);
void plvs_tsdf_destroy(plvs_tsdf* h);
int plvs_tsdf_reset(plvs_tsdf* h);                                  /* ChiselServer::Reset */
/* ChiselServer::SetDepthCameraInfo (ChiselServer.h:284) */
int plvs_tsdf_set_camera(plvs_tsdf* h, double fx, double fy, double cx, double cy, int w, int h_);

#define PLVS_TSDF_SCAN 0         /* Chisel::IntegrateDepthScan (Chisel.h:68): w=1, Carve() */
#define PLVS_TSDF_SCAN_COLOR 1   /* IntegrateDepthScanColorWithOneCameraModelBGR (Chisel.h:198) */

/* SetDepthPose + SetDepthImage[MemorySharing] (+SetColorImage) + IntegrateLastDepthImage(false).
 * With host buffers the call returns as soon as the (borrowed) buffers have been read -- the copy of scan k+1 overlaps
 * the kernels of scan k; statistics, read-outs and a pool-exhaustion error are delivered by the next call that needs the
 * finished map (plvs_tsdf_last_stats / download_blocks / export / reset).  Device-resident inputs (on_device=1) are not copied:
 * the call only enqueues, and the images must stay valid and unmodified until one of those waiting calls returns.
 * depth: float32 metres, w*h, row stride = w elements (DepthImage.h:54-74).  bgr: w*h*nch bytes
 * with `bgr_step` bytes per row (NULL unless mode==PLVS_TSDF_SCAN_COLOR).  Twc: 3x4 row-major
 * (rotation | translation), camera-to-world (src/PointCloudMapChisel.cc:147-154). */
int plvs_tsdf_integrate_depth(plvs_tsdf* h, const float* depth, int w, int h_,
                              const uint8_t* bgr, int bgr_step, int nch,
                              const float Twc[12], int mode, int on_device);

/* Step before the TSDF: the same scan from a 16-bit depth map (TUM / RealSense PNGs) without the host-side
 * `mImDepth.convertTo(mImDepth, CV_32F, mDepthMapFactor)` of src/Tracking.cc:1812-1813: the raw image goes over the bus (2 bytes per pixel),
 * `(float)d * depth_factor` runs on the device.  Host pointers only; otherwise as plvs_tsdf_integrate_depth (bgr rows are dense, bgr_stride is
 * unused as in chisel::ColorImage).  Two steps under the handle's lock: do not interleave with calls on the same handle from another thread. */
int plvs_tsdf_integrate_depth_u16(plvs_tsdf* h, const uint16_t* depth, int w, int h_, int stride_bytes, float depth_factor,
                                  const uint8_t* bgr, int bgr_stride, int nch, const float Twc[12], int mode);

/* Chisel::IntegratePointCloudWidthDepth (Thirdparty/open_chisel/src/Chisel.cpp:382-585) as reached through
 * ChiselServer::SetPointCloud + IntegrateLastPointCloud (PLVS's default Chisel route, src/PointCloudMapChisel.cc:100-131):
 * xyz = n camera-frame points (x,y,z float triples), rgb = n colour triples in [0,1] (r,g,b; NULL = no colour),
 * depth = the registered depth image used by the carve pass (NULL or carving off = no carve pass), Twc as above.
 * Host pointers.  Synchronous. */
int plvs_tsdf_integrate_cloud(plvs_tsdf* h, const float* xyz, const float* rgb, int n, const float* depth, int w, int h_,
                              const float Twc[12]);

typedef struct {
    int32_t n_blocks;            /* live chunks in the map */
    int32_t n_range;             /* chunks enumerated by GetChunkIDsIntersecting for the last scan */
    int32_t n_candidates;        /* chunks the voxel kernel actually visited */
    int32_t n_updated;           /* chunks with >=1 voxel update (needsUpdate==true) */
    int32_t n_new;               /* chunks created and kept */
    int32_t n_collected;         /* chunks created then garbage-collected */
    int32_t kernel_launches;
    int32_t pool_exhausted;      /* !=0 if max_blocks was hit (results incomplete) */
    int64_t total_updated;       /* since create/reset: sum of n_updated over all scans */
    int64_t total_candidates;    /* ... of chunks visited */
    int64_t total_integrations;  /* ... number of scans */
} plvs_tsdf_stats;
int plvs_tsdf_last_stats(const plvs_tsdf* h, plvs_tsdf_stats* out);
int plvs_tsdf_kernel_times(plvs_tsdf* h, float* ms, int32_t* launches, int reset);

/* Read-out: what PointCloudMapChisel::UpdateMap does after the integrations (src/PointCloudMapChisel.cc:233-250).
 * plvs_tsdf_update_meshes = ChiselServer::UpdateMesh -> Chisel::UpdateMeshes -> ChunkManager::RecomputeMesh (Thirdparty/open_chisel/src/
 * ChunkManager.cpp:116-172): marching cubes over every chunk (GenerateMesh :577-664), ColorizeMesh (:858-870) when the map has colour,
 * ComputeNormalsFromGradients (:838-856).  The meshes stay on the device; every chunk is re-meshed from the current voxels, which is what
 * the reference's dirty-set bookkeeping (27-neighbourhood of every updated chunk) amounts to.  *n_meshes = chunks with a non-empty mesh,
 * *n_verts = vertices over all of them (3 per triangle, not shared -- as chisel::Mesh stores them).
 * plvs_tsdf_get_meshes = reading ChunkManager::GetAllMeshes() (what ChiselServer::GetPointCloud walks, ChiselServer.cpp:872-975), in
 * (x,y,z) chunk-key order: keys[3*i], counts[i] = vertices of mesh i, then 3 floats per vertex for positions, normals and colours
 * (r,g,b in [0,1]; zeros when the map has no colour), concatenated in mesh order; vertex order inside a mesh is the reference's.
 * Any output may be NULL; vertex arrays may be device pointers (on_device != 0).  kfids are not carried. */
int plvs_tsdf_update_meshes(plvs_tsdf* h, int* n_meshes, long long* n_verts);
int plvs_tsdf_get_meshes(plvs_tsdf* h, int32_t* keys, int32_t* counts, int cap_meshes, float* verts, float* normals, float* colors,
                         long long cap_verts, int on_device);

/* Keyframe ids (DistVoxel::kfid, Thirdparty/open_chisel/include/open_chisel/DistVoxel.h:64-86; Mesh::kfids): the point-cloud route stamps every voxel it
 * integrates with the id of the point's keyframe (src/Chisel.cpp:470,534), Reset() clears it, and every mesh vertex carries the id of its cube's first
 * corner (ChunkManager.cpp:464-468) -- what ChunkManager::Deform keys on.  plvs_tsdf_integrate_cloud_kf = plvs_tsdf_integrate_cloud with
 * cloud.GetKfids(): kfids[n], or NULL for kfid_all on every point.  plvs_tsdf_download_kfid: per-voxel ids in the block order of
 * plvs_tsdf_download_blocks (kfid may be NULL to size).  plvs_tsdf_get_mesh_kfids: one id per vertex of the last plvs_tsdf_update_meshes.
 * The depth-scan route never writes ids; a voxel that route resets and re-integrates keeps the id of its last cloud here (0 in the reference). */
int plvs_tsdf_integrate_cloud_kf(plvs_tsdf* h, const float* xyz, const float* rgb, const uint32_t* kfids, uint32_t kfid_all, int n,
                                 const float* depth, int w, int h_, const float Twc[12]);
int plvs_tsdf_download_kfid(plvs_tsdf* h, uint32_t* kfid, int cap, int* n_out);
int plvs_tsdf_get_mesh_kfids(plvs_tsdf* h, uint32_t* kfids, long long cap_verts, int on_device);

/* ChiselServer::Deform(MapKfidRt&) -> ChunkManager::Deform (Thirdparty/open_chisel/src/ChunkManager.cpp:920-1062), the loop-closure correction of
 * PointCloudMapChisel (src/PointCloudMapChisel.cc:406-489): every known voxel whose keyframe id (plvs_tsdf_integrate_cloud_kf) is one of kfids[n] moves to
 * R * pos + t with Rt[12 * i] = (R row-major | t) of kfids[i]; voxels of other keyframes are dropped; voxels that land in one cell are folded in visiting
 * order (first: copy; later: DistVoxel::Integrate, SetKfid, ColorVoxel::Integrate(r,g,b,1)).  The reference visits its chunks in std::unordered_map order,
 * which no caller can know: chunk_order[3 * i] lists chunk ids to visit first (tests pass the compiled reference's own order), the remaining chunks follow
 * in (x,y,z) order -- n_order = 0 gives a deterministic result.  The meshes of the last plvs_tsdf_update_meshes move with their vertices.  The pool must
 * hold the old and the new map at once; otherwise PLVS_ENOMEM is returned and the map is left unchanged. */
int plvs_tsdf_deform(plvs_tsdf* h, const uint32_t* kfids, const float* Rt, int n, const int32_t* chunk_order, int n_order);

/* ChiselServer::IntegrateWorldPointCloud -> Chisel::IntegrateWorldPointCloudWithNormals (Thirdparty/open_chisel/src/Chisel.cpp:238-379), what
 * PointCloudMapChisel::LoadMap does with a saved map (src/PointCloudMapChisel.cc:527-549, after PCL has read the PLY file): every point updates the voxels
 * within 4 voxel sizes along its normal (u = (centre - point) . n, weight = w / (8 res)), stamps them with its keyframe id and colour.  xyz / normals: n
 * triples in the frame of Twc; rgb in [0,1] or NULL; kfids[n] or NULL for kfid_all. */
int plvs_tsdf_integrate_world_cloud(plvs_tsdf* h, const float* xyz, const float* rgb, const float* normals, const uint32_t* kfids, uint32_t kfid_all, int n,
                                    const float Twc[12]);

/* ChiselServer::SaveMesh -> Chisel::SaveAllMeshesToPLY (Thirdparty/open_chisel/src/Chisel.cpp:79-118, src/io/PLY.cpp:29-86): the vertices (and colours,
 * or NULL for a map without colour) that plvs_tsdf_get_meshes returned, as the reference's ASCII PLY.  Host I/O only. */
int plvs_mesh_save_ply(const char* path, const float* verts, const float* colors, long long n_verts);

/* The volumetric map on disk ("PLY save / load"): PointCloudMap<PointT>::WritePLY (src/PointCloudMap.cc:325-437) as
 * PointCloudMapChisel::SaveMap -> SaveTriangleMeshMap calls it (binary by default, one face per three consecutive vertices), byte for byte: vertex
 * properties x y z | red green blue | normal_x normal_y normal_z | label | kfid.  bgra: PCL's memory order (b, g, r, a), 4 bytes per point -- the binary form
 * stores its first three bytes under the names red, green, blue, exactly as the reference does.  Host I/O only. */
int plvs_map_save_ply(const char* path, const float* xyz, const uint8_t* bgra, const float* normals, const uint32_t* label, const uint32_t* kfid, long long n,
                      int is_mesh, int binary);
/* The reading half of PointCloudMap<PointT>::LoadMap (src/PointCloudMap.cc:466-503): vertex properties by name (ASCII or binary_little_endian; other
 * properties and the faces are skipped); rgb[3i..] = red, green, blue as named in the file.  *fields: 1 xyz, 2 rgb, 4 normals, 8 label, 16 kfid present.
 * PLVS_ECAP with *n_out set when the file holds more than cap points.  LoadMap's next steps -- InvertColors (swap red and blue) and
 * IntegrateWorldPointCloud -- are plvs_tsdf_integrate_world_cloud's caller's.  Host I/O only. */
int plvs_map_load_ply(const char* path, float* xyz, uint8_t* rgb, float* normals, uint32_t* label, uint32_t* kfid, long long cap, long long* n_out, int* fields);

/* read-out for tests/merge: chunk ids (x,y,z), per-voxel sdf / weight (4096 each, voxel index
 * (z*16+y)*16+x as Chunk.h:90-93) and rgba (r,g,b,colour-weight).  Any output may be NULL. */
int plvs_tsdf_download_blocks(plvs_tsdf* h, int32_t* keys, float* sdf, float* weight, uint8_t* rgba,
                              int cap, int* n_out);

/* Multi-GPU block merge (not in the reference).  export: pack every live block as
 * (key, w*sdf, w) into caller device buffers; merge: fold packed blocks into this map with the
 * commutative weighted sum.  The exchange itself (all-to-all by key owner) is NCCL plumbing done
 * by the caller (plvs_b200/parallel.py uses torch.distributed). */
int plvs_tsdf_export_packed(plvs_tsdf* h, int32_t* d_keys, float* d_wsdf, float* d_w, int cap, int* n_out);
int plvs_tsdf_merge_packed(plvs_tsdf* h, const int32_t* d_keys, const float* d_wsdf, const float* d_w, int n);
/* The same with the colour state: d_rgba[n][4096] = every voxel's ColorVoxel (r, g, b, colour weight; Thirdparty/open_chisel/include/open_chisel/
 * ColorVoxel.h:34-127) as stored.  The merge folds an incoming ColorVoxel in as ColorVoxel::Integrate would fold `weight` observations of that colour
 * (weighted mean truncated to a byte, weight saturating at 255), items of one block in list order.  d_rgba may be NULL (= the calls above);
 * any number of items per call. */
int plvs_tsdf_export_packed_rgba(plvs_tsdf* h, int32_t* d_keys, float* d_wsdf, float* d_w, uint32_t* d_rgba, int cap, int* n_out);
int plvs_tsdf_merge_packed_rgba(plvs_tsdf* h, const int32_t* d_keys, const float* d_wsdf, const float* d_w, const uint32_t* d_rgba, int n);

/* ------------------------------------------------------------------------------------------ */
/* Stream driver (host code only; plvs_b200/csrc/pipeline.cu)                                   */
/* ------------------------------------------------------------------------------------------ */
/* The reference's thread roles around the entry points above, as native threads: frame construction (plvs_orb_extract_batch on batches of
 * `batch` frames), Tracking (plvs_match_projection_last + plvs_match_projection_map per frame on the device-resident keypoints), LocalMapping
 * (plvs_match_triangulation against the previous frame) and PointCloudMapping (plvs_tsdf_integrate_depth per frame), src/System.cc:317-398.
 * It calls nothing but those public entry points; every call is synchronous for its stage, consecutive steps overlap as a pipeline.  The
 * caller-side inputs of the searches (SURVEY.md §8d "Matcher queries": projected map points, feature vectors, F12) are prepared per frame by the
 * caller, as Tracking / LocalMapping would.  bench.py times this function. */
typedef struct {
    int32_t valid;                    /* 0 = no searches for this frame (the first frame of a stream) */
    int32_t n_ql, n_qm;
    const plvs_last_query* ql;        /* SearchByProjection(Cur, Last): the last frame's map points, projected */
    const plvs_mp_query* qm;          /* SearchByProjection(F, vpMapPoints): the local map points in view */
    const float* uright;              /* mvuRight of this frame (host, one per keypoint) or NULL */
    plvs_featvec fv_cur, fv_last;     /* DBoW2 feature vectors of this frame and the previous one (host) */
    const uint8_t* has_cur;           /* per keypoint: already has a map point (this frame / the previous one) */
    const uint8_t* has_last;
    float F12[9], ep[2];              /* fundamental matrix and epipole of (this frame, previous frame) */
    plvs_frame_view last;             /* the previous frame (host view): KeyFrame 2 of SearchForTriangulation */
} plvs_pipeline_frame;

typedef struct {
    int32_t device, width, height, batch, n_steps, first_frame, cap;   /* cap: keypoint capacity per frame of the extractor outputs */
    int32_t inputs_on_device;         /* gray / depth / bgr are device pointers (inputs resident in HBM) */
    const uint8_t* gray;              /* frame f at gray + f*w*h */
    const float* depth;               /* depth + f*w*h (metres) */
    const uint8_t* bgr;               /* bgr + f*w*h*3, or NULL for a map without colour */
    const float* poses;               /* Twc of frame f at poses + 12*f */
    const plvs_pipeline_frame* frames;    /* indexed by absolute frame number */
    plvs_frame_view view_template;    /* bounds, grid cell sizes, scale factors, level sigmas, bf shared by the stream's frames */
    float th_last, th_map, nnratio_map;   /* 15, 3, 0.8 in Tracking */
    void* flush_buf; size_t flush_bytes;  /* optional device buffer rewritten at the start of every step (evicts L2 between steps) */
} plvs_pipeline_job;

typedef struct {
    int64_t keypoints, matches;
    double wall_s, busy_extract_s, busy_track_s, busy_tri_s, busy_map_s;
} plvs_pipeline_stats;

/* steps first_frame + s*batch .. for s in [0, n_steps); ex[s % n_ex] extracts step s (n_ex even, 2..8: n_ex / 2 frame-construction threads take
 * alternate steps, so with n_ex = 4 two batches are in extraction while Tracking / LocalMapping read a third).  Returns when every stage has drained. */
int plvs_pipeline_run(plvs_orb* const* ex, int n_ex, plvs_match* m_track, plvs_match* m_tri, plvs_tsdf* tsdf, const plvs_pipeline_job* job,
                      plvs_pipeline_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* PLVS_B200_H_ */
