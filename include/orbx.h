/* orbx.h — C ABI of the B200-native ORB feature-extraction front-end.
 *
 * Drop-in boundary for TC2LI_SLAM::ORBextractor (reference: slam_framework/include/ORBextractor.h:52-118,
 * slam_framework/src/ORBextractor.cc:383-443, 1060-1168).  The reference has no FFI layer of its own: the seam is
 * the C++ class, whose replacement (tc2li_slam_b200/compat/ORBextractor.h) is a header-only shim over this ABI.
 * Plain pointers and sizes only; no C++, CUDA or torch types cross this boundary.
 *
 * Threading (mirrors ORBextractor.h:92 / Frame.cc:139-142): calls on ONE handle are serialised by the caller;
 * calls on DIFFERENT handles may run concurrently from different host threads (each handle owns its CUDA stream
 * and device buffers).  All functions return ORBX_OK (0) or a negative orbx_status unless stated otherwise.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_MAX_LEVELS 16
#define ORBX_MAX_DIM 4096 /* candidate coordinates are packed in 12 bits */

typedef enum orbx_status {
    ORBX_OK = 0,
    ORBX_ERR_EMPTY = -1,     /* empty image: the value ORBextractor::operator() returns (ORBextractor.cc:1063-1064) */
    ORBX_ERR_TOO_SMALL = -2, /* some pyramid level has no 35-px FAST cell (the reference divides by zero there) */
    ORBX_ERR_ARG = -3,
    ORBX_ERR_CUDA = -4,
    ORBX_ERR_CAPACITY = -5, /* caller's output capacity < orbx_max_keypoints() */
    ORBX_ERR_TOO_LARGE = -6 /* width or height > ORBX_MAX_DIM, or per-level quota too large for one CTA's shared memory */
} orbx_status;

/* Layout-identical to cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id.
 * class_id stays -1, response is the FAST score, size = (float)(int)(31 * mvScaleFactor[octave]) (ORBextractor.cc:854-864). */
typedef struct orbx_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

/* ------------------------------------------------------------------------------------------------------------------
 * Single-image handle == one ORBextractor instance
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct orbx_extractor orbx_extractor;

/* replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) (ORBextractor.cc:383-443);
 * `device` = CUDA device ordinal.  Fails with ORBX_ERR_CUDA when no usable GPU exists: there is no CPU fallback. */
int orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device, orbx_extractor** out);
void orbx_destroy(orbx_extractor* h);

/* replace GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (ORBextractor.h:70-90); `out` receives nlevels floats.  features_per_level mirrors the
 * protected mnFeaturesPerLevel (ORBextractor.cc:408-419) for tests. */
int orbx_get_levels(const orbx_extractor* h);
float orbx_get_scale_factor(const orbx_extractor* h);
int orbx_get_scale_factors(const orbx_extractor* h, float* out);
int orbx_get_inverse_scale_factors(const orbx_extractor* h, float* out);
int orbx_get_scale_sigma_squares(const orbx_extractor* h, float* out);
int orbx_get_inverse_scale_sigma_squares(const orbx_extractor* h, float* out);
int orbx_get_features_per_level(const orbx_extractor* h, int* out);

/* Upper bound of keypoints one call can emit: sum over levels of (quota + 3, at least 4 * initial nodes);
 * the reference overshoots nfeatures by up to 2 per level (ORBextractor.cc:720-721). */
int orbx_max_keypoints(const orbx_extractor* h);

/* GaussianBlur 8.8 fixed-point taps (default {18,34,48,56,48,34,18} = OpenCV >= 4.5; OpenCV 4.2 rounded each tap
 * independently).  Replaces nothing in the reference: it pins the un-vendored OpenCV version. */
int orbx_set_blur_taps(orbx_extractor* h, const int taps[7]);

/* replaces int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea) (ORBextractor.cc:1060-1141).
 * image: 8-bit gray, `stride` bytes per row, HOST memory.  lap0/lap1 = vLappingArea[0..1].
 * keypoints[capacity], descriptors[capacity*32] are HOST buffers, capacity >= orbx_max_keypoints().
 * *n_keypoints = N, *mono_index = the reference's return value (keypoints outside the lapping area come first, the
 * ones inside are filled from the back).  Returns ORBX_ERR_EMPTY (-1) for a null/empty image like the reference. */
int orbx_extract(orbx_extractor* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                 orbx_keypoint* keypoints, uint8_t* descriptors, int capacity, int* n_keypoints, int* mono_index);

/* replaces the two std::thread extractions of Frame::Frame (Frame.cc:139-142, 1090-1094).  Handles with identical parameters (the
 * reference builds mpORBextractorLeft / Right from the same settings) share ONE launch sequence over a batch of two images — the
 * sequence is latency bound, so the pair costs what one image costs; otherwise both images are issued on the two handles' own CUDA
 * streams before either is waited for.  lap_left / lap_right = {x0, x1} as Frame::ExtractORB builds them (Frame.cc:445-455).
 * Either way orbx_get_level / orbx_stereo_match / orbx_lapping_match on the handles see this call's results. */
int orbx_extract_stereo(orbx_extractor* left, orbx_extractor* right, const uint8_t* image_left, const uint8_t* image_right,
                        int width, int height, size_t stride_left, size_t stride_right, const int lap_left[2], const int lap_right[2],
                        orbx_keypoint* kps_left, uint8_t* desc_left, int cap_left, int* n_left, int* mono_left,
                        orbx_keypoint* kps_right, uint8_t* desc_right, int cap_right, int* n_right, int* mono_right);

/* ------------------------------------------------------------------------------------------------------------------
 * The step BEFORE the extractor in the caller: Tracking::GrabImageStereoLidar (Tracking.cc:1646-1671, and the same block in
 * the other Grab* entry points) converts 3- and 4-channel inputs with cvtColor(COLOR_RGB2GRAY | COLOR_BGR2GRAY |
 * COLOR_RGBA2GRAY | COLOR_BGRA2GRAY), chosen by Camera.RGB (mbRGB), before it constructs the Frame.  These entry points take
 * the interleaved 8-bit colour image as it is (channels = 3 or 4, rgb != 0 <=> mbRGB: the first channel is red), convert on
 * the device with OpenCV 4's arithmetic, gray = (R*9798 + G*19235 + B*3735 + 2^14) >> 15, bit-identical to cv::cvtColor, and
 * extract.  The grey image the reference keeps (Tracking::mImGray) is level 0: orbx_get_level(h, 0, dst, stride, 0).
 * ---------------------------------------------------------------------------------------------------------------- */
int orbx_extract_color(orbx_extractor* h, const uint8_t* image, int width, int height, size_t stride, int channels, int rgb, int lap0, int lap1,
                       orbx_keypoint* keypoints, uint8_t* descriptors, int capacity, int* n_keypoints, int* mono_index);
int orbx_extract_stereo_color(orbx_extractor* left, orbx_extractor* right, const uint8_t* image_left, const uint8_t* image_right, int width, int height,
                              size_t stride_left, size_t stride_right, int channels, int rgb, const int lapping_left[2], const int lapping_right[2],
                              orbx_keypoint* kps_left, uint8_t* desc_left, int cap_left, int* n_left, int* mono_left,
                              orbx_keypoint* kps_right, uint8_t* desc_right, int cap_right, int* n_right, int* mono_right);

/* ------------------------------------------------------------------------------------------------------------------
 * One step further up the caller: System::TrackStereoLidar (System.cc:238-248) rectifies both images with
 * cv::remap(im, imToFeed, M1, M2, cv::INTER_LINEAR), M1/M2 = the CV_32FC1 maps Settings builds with
 * cv::initUndistortRectifyMap (Settings.cc:523-526), when Settings::needToRectify().  orbx_set_rectification hands this
 * extractor the maps of ITS camera (M1l/M2l for the left, M1r/M2r for the right instance): width x height floats each,
 * sampling raw images of src_width x src_height.  From then on orbx_extract / orbx_extract_stereo / orbx_extract_color /
 * orbx_extract_stereo_color take the RAW image (of the raw size); remap (per channel, constant-0 border, OpenCV's 5-bit
 * fixed-point coordinates and 15-bit weights), the grey conversion and the extraction run on the device in the reference's
 * order, bit-identical to cv::remap + cv::cvtColor.  NULL maps switch rectification off.
 * ---------------------------------------------------------------------------------------------------------------- */
int orbx_set_rectification(orbx_extractor* h, const float* map_x, const float* map_y, int width, int height, size_t map_stride_bytes, int src_width,
                           int src_height);
/* the other branch of the same `if`: cv::resize(im, imToFeed, settings_->newImSize()) (System.cc:250-252, Settings::needToResize()).  After this call the
 * orbx_extract* entry points take the RAW src_width x src_height image and extract from its width x height resize (INTER_LINEAR, bit-identical to
 * cv::resize; colour input is converted afterwards as cv::cvtColor does).  It replaces a rectification and vice versa; width = height = 0 switches it off. */
int orbx_set_input_resize(orbx_extractor* h, int width, int height, int src_width, int src_height);

/* ------------------------------------------------------------------------------------------------------------------
 * Next step of the caller: replaces Frame::ComputeStereoMatches (Frame.cc:841-1011) and the
 * ORBmatcher::DescriptorDistance calls inside it (ORBmatcher.cc:2067-2083) for the stereo pair last extracted by `left` and
 * `right`; keypoints, descriptors and both pyramids stay on the device.  mb, mbf = Frame::mb / Frame::mbf (baseline in
 * metres, baseline * fx).  uright[capacity] / depth[capacity] receive Frame::mvuRight / Frame::mvDepth for the *n_left
 * left keypoints (-1 = no match).  Bit-identical to the reference.
 * ---------------------------------------------------------------------------------------------------------------- */
int orbx_stereo_match(orbx_extractor* left, orbx_extractor* right, float mb, float mbf, float* uright, float* depth, int capacity, int* n_left);

/* ------------------------------------------------------------------------------------------------------------------
 * Next step of the fisheye (KannalaBrandt8) stereo caller: the descriptor stage of
 * Frame::ComputeStereoFishEyeMatches (Frame.cc:1162-1187) — `BFmatcher.knnMatch(stereoDescLeft, stereoDescRight, matches, 2)`
 * with BFmatcher = cv::BFMatcher(cv::NORM_HAMMING) (Frame.cc:52) over the lapping-area tails [monoLeft, Nleft) /
 * [monoRight, Nright) that operator() leaves at the back of both outputs (ORBextractor.cc:1126-1135), and Lowe's gate
 * `m[0].distance < m[1].distance * 0.7` (Frame.cc:1187).  For query q (left keypoint mono_left + q) knn[q] holds the two
 * nearest train rows (right keypoint mono_right + train_idx) in cv::DMatch order, -1 where the right tail has fewer rows;
 * ratio_ok[q] != 0 iff the reference would go on to KannalaBrandt8::TriangulateMatches for that pair (Frame.cc:1188-1192,
 * which stays on the host).  Bit-identical to OpenCV's brute-force matcher, ties included.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct orbx_knn2 {
    int train_idx0, distance0; /* cv::DMatch{queryIdx = q, trainIdx, distance} of the nearest train row */
    int train_idx1, distance1; /* ... of the second nearest */
} orbx_knn2;
int orbx_lapping_match(orbx_extractor* left, orbx_extractor* right, orbx_knn2* knn, uint8_t* ratio_ok, int capacity, int* n_query, int* n_train);

/* replace the public member mvImagePyramid (ORBextractor.h:92), read by Frame::ComputeStereoMatches
 * (Frame.cc:848,938,948,953): size and (lazy) device->host copy of level `level` of the last extracted image.
 * border = 0 copies w x h pixels to dst; border = 19 writes the (w+38) x (h+38) BORDER_REFLECT_101 image the
 * reference allocates (ORBextractor.cc:1149-1165), dst pointing at its first byte. */
int orbx_get_level_size(const orbx_extractor* h, int level, int* width, int* height);
int orbx_get_level(orbx_extractor* h, int level, uint8_t* dst, size_t dst_stride, int border);

/* ------------------------------------------------------------------------------------------------------------------
 * Batched extractor: n images of one shape per launch set (frames x {left,right} are just consecutive images).
 * This is the throughput path (BASELINE.json configs[4]); the single-image handle above is a batch of one.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct orbx_batch orbx_batch;

int orbx_batch_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device, int width, int height,
                      int max_images, orbx_batch** out);
void orbx_batch_destroy(orbx_batch* b);
int orbx_batch_max_keypoints(const orbx_batch* b); /* per image: K */
int orbx_batch_max_images(const orbx_batch* b);
int orbx_batch_set_blur_taps(orbx_batch* b, const int taps[7]);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores the handle's own stream) */
int orbx_batch_set_stream(orbx_batch* b, void* cuda_stream);
void* orbx_batch_get_stream(const orbx_batch* b);

/* inputs: host images (async H2D; pinned memory makes it overlap), device images (D2D), or the on-device synthetic
 * generator orbx-synth-v1.  Image i of a stereo batch: frame = i/2, right image when i is odd
 * (sampled at x+16, noise seed ^ 0x5bd1e995); seed = seed0 + frame. */
int orbx_batch_upload(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int first_image, int n_images);
/* colour variant (see orbx_extract_color): interleaved 8-bit images, row_stride >= width * channels; grey conversion on the device */
int orbx_batch_upload_color(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int channels, int rgb, int first_image, int n_images);
/* rectification variant (see orbx_set_rectification): maps of width x height floats for ALL images of the handle; then raw images of
 * src_width x src_height with 1, 3 or 4 channels are remapped (+ converted to grey) into level 0 */
int orbx_batch_set_rectification(orbx_batch* b, const float* map_x, const float* map_y, size_t map_stride_bytes, int src_width, int src_height);
/* resize variant: System::TrackStereoLidar's cv::resize(im, imToFeed, settings_->newImSize()) (System.cc:250-252, Settings::needToResize())
 * on the device: after this call orbx_batch_upload_raw takes raw images of src_width x src_height (1, 3 or 4 channels) and resizes them
 * (INTER_LINEAR, bit-identical to cv::resize; colour input is then converted as cv::cvtColor does) into level 0.  It replaces a
 * rectification set earlier and vice versa; 0, 0 switches it off. */
int orbx_batch_set_input_resize(orbx_batch* b, int src_width, int src_height);
int orbx_batch_upload_raw(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int channels, int rgb, int first_image, int n_images);
int orbx_batch_upload_device(orbx_batch* b, const uint8_t* dev_images, size_t row_stride, size_t image_stride, int first_image, int n_images);
int orbx_batch_synth(orbx_batch* b, uint32_t seed0, int first_image, int n_images, int stereo);
/* level-0 images [first, first+n) back to a dense host array (synchronous) */
int orbx_batch_download_images(orbx_batch* b, int first_image, int n_images, uint8_t* host_images, size_t row_stride, size_t image_stride);

/* all kernels of the hot path for images [0, n_images), asynchronous on the handle's stream.
 * lapping: n_images x {lap0, lap1} (host ints) or NULL for {0,0} everywhere. */
int orbx_batch_run(orbx_batch* b, int n_images, const int* lapping);
/* async D2H of the results + stream synchronisation.  keypoints [n_images][K], descriptors [n_images][K][32],
 * counts [n_images], mono_index [n_images]; any pointer may be NULL to skip that array. */
int orbx_batch_download(orbx_batch* b, int n_images, orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index);
/* batched Frame::ComputeStereoMatches: image 2f = left, 2f+1 = right image of frame f (after orbx_batch_run of those images).
 * orbx_batch_download_stereo: uright / depth [n_frames][K] floats (K = orbx_batch_max_keypoints), -1 = no match; synchronises. */
int orbx_batch_stereo_match(orbx_batch* b, int n_frames, float mb, float mbf);
int orbx_batch_download_stereo(orbx_batch* b, int n_frames, float* uright, float* depth);
/* batched descriptor stage of Frame::ComputeStereoFishEyeMatches (see orbx_lapping_match): image 2f = left, 2f+1 = right.
 * orbx_batch_download_lapping: knn / ratio_ok [n_frames][K]; row f holds the N_left - mono_left queries of frame f; synchronises. */
int orbx_batch_lapping_match(orbx_batch* b, int n_frames);
int orbx_batch_download_lapping(orbx_batch* b, int n_frames, orbx_knn2* knn, uint8_t* ratio_ok);
/* The feature grid of a Frame on the resident keypoints (after orbx_batch_run): Frame::AssignFeaturesToGrid + Frame::PosInGrid
 * (Frame.cc:412-443, 751-762) for the single-grid case (Nleft == -1; rectified input, so mvKeysUn == mvKeys, Frame.cc:777-782).
 * Bounds = Frame::mnMinX / mnMaxX / mnMinY / mnMaxY (Frame::ComputeImageBounds, Frame.cc:812-839: 0, cols, 0, rows without
 * distortion); the cell sizes follow Frame.cc:179-180.  Frame::mGrid[i][j] comes back as CSR: cell c = i * ORBX_GRID_ROWS + j
 * holds items[cell_start[c] .. cell_start[c + 1]) = keypoint indices in the reference's push_back order.  Asynchronous.
 * orbx_batch_download_grid: cell_start [n_images][ORBX_GRID_CELLS + 1], items [n_images][K]; either may be NULL; synchronises. */
#define ORBX_GRID_COLS 64 /* FRAME_GRID_COLS, Frame.h:60 */
#define ORBX_GRID_ROWS 48 /* FRAME_GRID_ROWS, Frame.h:59 */
#define ORBX_GRID_CELLS (ORBX_GRID_COLS * ORBX_GRID_ROWS)
/* Frame::UndistortKeyPoints (Frame.cc:777-810) on the resident keypoints of images [0, n_images): cv::undistortPoints(mat, mat, toK(), mDistCoef, cv::Mat(), mK),
 * bit-identical to OpenCV (five fixed-point iterations in double).  camera / new_camera = {fx, fy, cx, cy} of Pinhole::toK() / Frame::mK; dist = mDistCoef
 * (n_dist = 4, 5, 8 or 12).  dist == NULL or dist[0] == 0: mvKeysUn = mvKeys, as the reference's early return.  Asynchronous.  A grid built afterwards
 * (orbx_batch_assign_grid) and the single-grid searches on it use the undistorted coordinates (the caller passes the bounds of Frame::ComputeImageBounds);
 * two-grid (fisheye) frames use mvKeys / mvKeysRight as in the reference.  The next orbx_batch_run* resets it.
 * orbx_batch_download_undistorted: xy [n_images][K][2] floats = mvKeysUn[i].pt; synchronises. */
int orbx_batch_undistort_keypoints(orbx_batch* b, int n_images, const float camera[4], const float* dist, int n_dist, const float new_camera[4]);
int orbx_batch_download_undistorted(orbx_batch* b, int n_images, float* xy);
int orbx_batch_assign_grid(orbx_batch* b, int n_images, float min_x, float max_x, float min_y, float max_y);
int orbx_batch_download_grid(orbx_batch* b, int n_images, int* cell_start, int* items);
/* Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (Frame.cc:687-749) for n_queries windows against the grids of
 * orbx_batch_assign_grid (host arrays in, host arrays out; synchronises): counts[q] = the size of the vector the reference returns,
 * indices[q * capacity ..] = its first min(counts[q], capacity) entries in the reference's order. */
typedef struct orbx_area_query {
    int image;
    float x, y, r;
    int min_level, max_level; /* the reference's defaults are -1, -1 */
} orbx_area_query;
int orbx_batch_features_in_area(orbx_batch* b, const orbx_area_query* queries, int n_queries, int capacity, int* counts, int* indices);
/* The search of ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1685-1895) from the
 * point where a map point has been projected (ORBmatcher.cc:1729 on), single-grid case, against image `image` of the handle (after
 * orbx_batch_run and orbx_batch_assign_grid; use_uright = 1 also needs orbx_batch_stereo_match, whose mvuRight the gate of
 * ORBmatcher.cc:1760-1766 reads).  The caller keeps the pose algebra (Tcw * x3Dw, project, the bounds test, the bForward / bBackward decision)
 * and sends one query per surviving map point, in LastFrame order:
 *   u, v = uv(0), uv(1);  radius = th * mvScaleFactors[nLastOctave];  ur = uv(0) - mbf * invzc;
 *   min_level / max_level = the GetFeaturesInArea level window of its branch: (o, -1) forward, (0, o) backward, (o - 1, o + 1) otherwise;
 *   angle = kpLF.angle;  blocks = pMP->Observations() > 0;  desc = pMP->GetDescriptor().
 * match[k] (k < K = orbx_batch_max_keypoints) = index of the query whose map point CurrentFrame.mvpMapPoints[k] ends up with, -1 = NULL
 * (CurrentFrame.mvpMapPoints starts all-NULL, as Tracking::TrackWithMotionModel fills it before the call); *n_matches = the value the
 * reference returns.  th_high = ORBmatcher::TH_HIGH (100); check_orientation = ORBmatcher::mbCheckOrientation.  Host arrays; synchronises. */
typedef struct orbx_proj_query {
    float u, v, radius, ur;
    int min_level, max_level;
    float angle;
    int blocks;
    uint8_t desc[32];
} orbx_proj_query;
int orbx_batch_search_by_projection(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int use_uright, int check_orientation, int th_high,
                                    int* match, int* n_matches);
/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:52-222), the
 * local-map search of Tracking::SearchLocalPoints (Tracking.cc:3282), single-grid case.  One query per map point that passes the early tests
 * (ORBmatcher.cc:61-70: mbTrackInView, the far-point test, isBad), in vpMapPoints order:
 *   u, v = mTrackProjX, mTrackProjY;  radius = r * F.mvScaleFactors[nPredictedLevel], r = RadiusByViewingCos(mTrackViewCos) (* th if th != 1);
 *   ur = mTrackProjXR;  min_level, max_level = nPredictedLevel - 1, nPredictedLevel;  blocks = Observations() > 0;  desc = GetDescriptor()  (angle unused).
 * taken[k] != 0 (K host bytes, or NULL for none) where F.mvpMapPoints[k] holds a point with observations before the search.  nn_ratio =
 * ORBmatcher::mfNNratio.  match[k] = index of the query this search writes into F.mvpMapPoints[k], -1 = left as it was; *n_matches = the
 * reference's return value.  Same prerequisites as orbx_batch_search_by_projection; synchronises. */
int orbx_batch_search_local_points(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int use_uright, float nn_ratio, int th_high,
                                   const uint8_t* taken, int* match, int* n_matches);
/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1898-2019),
 * the search of Tracking::Relocalization, from the point where a map point of the keyframe has been projected and its level predicted
 * (ORBmatcher.cc:1947 on).  One query per surviving point, in pKF->GetMapPointMatches() order: u, v = uv; radius = th * mvScaleFactors[nPredictedLevel];
 * min_level, max_level = nPredictedLevel - 1, nPredictedLevel + 1; angle = pKF->mvKeysUn[i].angle; blocks = 1 (here every assigned keypoint is
 * skipped, ORBmatcher.cc:1962-1963); desc = GetDescriptor(); ur unused.  taken[k] != 0 (K host bytes or NULL) where CurrentFrame.mvpMapPoints[k]
 * is non-NULL before the search.  orb_dist = ORBdist.  match / n_matches as for orbx_batch_search_by_projection; synchronises. */
int orbx_batch_search_keyframe_points(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int check_orientation, int orb_dist,
                                      const uint8_t* taken, int* match, int* n_matches);
/* The same two searches for a TWO-GRID frame (fisheye stereo, Frame::Nleft != -1): ORBmatcher.cc:1803-1868 (frame to frame) and :153-218 (local map).
 * The frame is image_left + image_right of the handle (each with its own grid from orbx_batch_assign_grid = Frame::mGrid / mGridRight, Frame.cc:428-441);
 * mvpMapPoints[0 .. Nleft) is match_left, mvpMapPoints[Nleft ..) is match_right (K entries each).  A map point is sent as up to two queries IN A ROW,
 * told apart by side[i]:  0 = its search in the left image (u, v = its projection there);  1 = its search in the right image, directly after its
 * side-0 query (u, v = the projection into the right camera: GetRelativePoseTrl() * x3Dc resp. mTrackProjXR / mTrackProjYR, radius from mTrackViewCosR
 * without the th factor, levels from mnTrackScaleLevelR);  2 = a right-image search of a point without a left part (local map: mbTrackInViewR only).
 * The device replays what the reference's `continue`s do: an empty left window cancels the right search of that point in the frame-to-frame search
 * (:1744-1745), a failed ratio test in the local-map search (:133-134).  The right-coordinate gate does not exist for such frames (ur unused).
 * left_to_right / right_to_left = Frame::mvLeftToRightMatch / mvRightToLeftMatch (K host ints each, -1 = none, or NULL): an accepted local-map match also
 * takes the partner keypoint and counts twice (:139-143, :205-209).  match values are indices into `queries`.  Synchronise. */
int orbx_batch_search_by_projection_stereo(orbx_batch* b, int image_left, int image_right, const orbx_proj_query* queries, const uint8_t* side, int n_queries,
                                           int check_orientation, int th_high, int* match_left, int* match_right, int* n_matches);
int orbx_batch_search_local_points_stereo(orbx_batch* b, int image_left, int image_right, const orbx_proj_query* queries, const uint8_t* side, int n_queries,
                                          float nn_ratio, int th_high, const uint8_t* taken_left, const uint8_t* taken_right, const int* left_to_right,
                                          const int* right_to_left, int* match_left, int* match_right, int* n_matches);
/* ------------------------------------------------------------------------------------------------------------------
 * Bag of words: Frame::ComputeBoW (Frame.cc:768-775; KeyFrame::ComputeBoW is the same call on stored descriptors)
 *     mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
 * = DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, v, fv, levelsup)
 * (slam_framework/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1276, FORB::distance FORB.cpp:81-100, BowVector.cpp:34-84).
 * orbx_vocabulary = the resident ORBVocabulary (the tree, ~35 MB for ORBvoc.txt): orbx_vocabulary_load_text reads the text format
 * of ORBvoc.txt like ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1350-1433, called once by System's constructor);
 * orbx_vocabulary_create takes the same tree as flat arrays (node 0 = root; parent[n] < n; is_word[n] marks the words; descriptors
 * [n][32]; weights[n]).  Results, per image / call, bit-identical to DBoW2's containers:
 *   BowVector     : words[0..n_words) ascending WordIds, values[] their WordValues (doubles, normalised as the scoring type asks)
 *   FeatureVector : nodes[0..n_nodes) ascending NodeIds at level L - levelsup, features[node_start[j] .. node_start[j + 1]) the
 *                   feature indices of nodes[j] in ascending order
 * Per-image capacity of every array = K (orbx_batch_max_keypoints), node_start K + 1.  At most ORBX_BOW_MAX_FEATURES features per image.
 * ---------------------------------------------------------------------------------------------------------------- */
#define ORBX_BOW_MAX_FEATURES 8192
typedef struct orbx_vocabulary orbx_vocabulary;
int orbx_vocabulary_load_text(int device, const char* path, orbx_vocabulary** out);
int orbx_vocabulary_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* is_word,
                           const uint8_t* descriptors, const double* weights, orbx_vocabulary** out);
void orbx_vocabulary_destroy(orbx_vocabulary* v);
int orbx_vocabulary_info(const orbx_vocabulary* v, int out[6]); /* {k, L, scoring, weighting, nodes incl. root, words} */
/* transform() of n host descriptors [n][32] (capacity of every output array: n, node_start n + 1); synchronous.  Uses a stream and scratch buffers of the
 * vocabulary handle: one call at a time per handle (orbx_batch_compute_bow only reads the vocabulary and may run concurrently on many batches). */
int orbx_vocabulary_transform(orbx_vocabulary* v, const uint8_t* descriptors, int n, int levelsup, uint32_t* words, double* values, int* n_words,
                              uint32_t* nodes, int* node_start, uint32_t* features, int* n_nodes);
/* transform() of the descriptors resident after orbx_batch_run, images [0, n_images); asynchronous on the handle's stream.
 * orbx_batch_download_bow: words / nodes / features [n_images][K], values [n_images][K] doubles, node_start [n_images][K + 1],
 * n_words / n_nodes [n_images]; any pointer may be NULL; synchronises. */
int orbx_batch_compute_bow(orbx_batch* b, const orbx_vocabulary* v, int n_images, int levelsup);
int orbx_batch_download_bow(orbx_batch* b, int n_images, uint32_t* words, double* values, int* n_words, uint32_t* nodes, int* node_start,
                            uint32_t* features, int* n_nodes);
int orbx_batch_sync(orbx_batch* b);
/* the same on a sub-range [first, first+n) of the handle's images (lapping: n x 2 ints for that range): lets a caller keep
 * many frames resident and process them in groups small enough for the 126 MB L2 */
int orbx_batch_run_range(orbx_batch* b, int first, int n, const int* lapping);
/* images [0, n) in chunks of `chunk` images issued round-robin over the handle's internal streams: the latency-bound stages of
 * one chunk overlap the issue-bound stages of another.  Asynchronous like orbx_batch_run; the handle's stream waits for all chunks. */
int orbx_batch_run_pipelined(orbx_batch* b, int n, const int* lapping, int chunk);
/* D2H of a sub-range without waiting (pinned destinations make it truly asynchronous); orbx_batch_check() synchronises the
 * stream and reports device-side capacity errors */
int orbx_batch_download_async(orbx_batch* b, int first, int n, orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index);
int orbx_batch_check(orbx_batch* b);

/* CUDA events on the handles' streams (device-side timing, ordering between handles of a pipelined caller) */
void* orbx_event_create(const orbx_batch* b);   /* on the batch's device */
void orbx_event_destroy(void* ev);
int orbx_event_record(void* ev, orbx_batch* b);
int orbx_event_wait(orbx_batch* b, void* ev);
int orbx_event_synchronize(void* ev);
int orbx_event_elapsed_ms(void* start, void* stop, float* ms);
int orbx_batch_get_level_size(const orbx_batch* b, int level, int* width, int* height);
int orbx_batch_get_level(orbx_batch* b, int image, int level, uint8_t* dst, size_t dst_stride, int border);

/* per-stage device timing of the last orbx_batch_run (CUDA events on the launching stream), enabled by
 * orbx_batch_enable_timing(b, 1).  Stages: 0 pyramid, 1 FAST, 2 blur, 3 octree select, 4 assemble, 5 orient+describe. */
#define ORBX_NUM_STAGES 6
int orbx_batch_enable_timing(orbx_batch* b, int on);
int orbx_batch_stage_ms(orbx_batch* b, float ms[ORBX_NUM_STAGES]);
int orbx_batch_kernel_launches(const orbx_batch* b); /* kernels launched by one orbx_batch_run */

/* stage dumps for the parity tests (synchronise the stream).  Candidates come back in the reference's order
 * (cell row-major, raster inside the cell, after the per-cell iniTh/minTh decision): x, y relative to (16,16), response. */
int orbx_batch_debug_blurred(orbx_batch* b, int image, int level, uint8_t* dst, size_t dst_stride);
int orbx_batch_debug_candidates(orbx_batch* b, int image, int level, float* xyr, int capacity, int* n);
int orbx_batch_debug_selected(orbx_batch* b, int image, int level, float* xyr, int capacity, int* n);

int orbx_batch_debug_fallback_cells(orbx_batch* b, int image, int level, int* n_fallback);
/* the batch that holds the results of the handle's last call (NULL before the first call): the batch-of-one of the handle after a single-image call; after
 * orbx_extract_stereo on two handles with identical parameters the batch of TWO that served the pair (image 0 = left, image 1 = right: the layout the
 * orbx_batch_* pair functions expect), returned for both handles */
orbx_batch* orbx_extractor_batch(orbx_extractor* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU pool: the product form of "frames are independent, shard them by frame over the GPUs of one box, no collective"
 * (SURVEY.md 8e).  It scales out what every Frame constructor does for one stereo pair — two extraction threads, one per
 * image (Frame.cc:139-142, Frame::ExtractORB Frame.cc:445-455): here one host thread per GPU, each driving `handles_per_device`
 * batch handles (own streams) over chunks of `chunk_images` images, so that the upload of one chunk, the kernels of the next
 * and the download of a third overlap.  Chunk c of a call goes to device c mod n_devices; the results land in the caller's
 * arrays at the images' own indices (frame order).  No torch, no NCCL, no device-to-device traffic.
 *   images      : n_images x height rows of row_stride bytes, image_stride bytes apart (pinned memory from orbx_alloc_pinned
 *                 keeps the copies asynchronous; pageable memory works, slower)
 *   lapping     : n_images x 2 ints (vLappingArea of every image, ORBextractor.cc:1126-1135) or NULL
 *   keypoints   : n_images x K (K = orbx_pool_max_keypoints), descriptors n_images x K x 32, counts / mono_index n_images;
 *                 any of the four may be NULL
 * orbx_pool_extract returns when everything is on the host (ORBX_ERR_CAPACITY etc. as for the batch handle). */
typedef struct orbx_pool orbx_pool;
int orbx_pool_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, const int* devices, int n_devices, int width,
                     int height, int chunk_images, int handles_per_device, orbx_pool** out);
void orbx_pool_destroy(orbx_pool* p);
int orbx_pool_max_keypoints(const orbx_pool* p);
int orbx_pool_devices(const orbx_pool* p);
int orbx_pool_extract(orbx_pool* p, const uint8_t* images, size_t row_stride, size_t image_stride, int n_images, const int* lapping,
                      orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index);
/* device time of the last orbx_pool_extract: per device, first upload enqueued -> last download finished (CUDA events on the
 * handles' streams), ms[n_devices]; the call's throughput is n_images / max(ms) */
int orbx_pool_last_ms(const orbx_pool* p, float* ms);
int orbx_pool_kernel_launches(const orbx_pool* p);   /* kernels launched by the last orbx_pool_extract, all devices */

/* page-locked host memory (portable: pinned for every device), so that the uploads / downloads above overlap with kernels of other handles */
void* orbx_alloc_pinned(size_t bytes);
void orbx_free_pinned(void* p);

/* ------------------------------------------------------------------------------------------------------------------
 * Host-only helpers (no GPU needed): the constants ORBextractor's constructor derives (ORBextractor.cc:383-443) and the
 * per-level geometry (level size :1147-1148, FAST grid :771-777, quadtree roots :533, keypoint capacity).
 * out[8] = {w, h, nCols, nRows, wCell, hCell, nIni, capacity}.
 * ---------------------------------------------------------------------------------------------------------------- */
int orbx_tables(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, float* scale_factors, float* inv_scale_factors,
                float* level_sigma2, float* inv_level_sigma2, int* features_per_level, int umax[16]);
int orbx_level_geometry(int nfeatures, float scale_factor, int nlevels, int width, int height, int level, int out[8]);

/* test hooks for the exactness-critical routines.  orbx_debug_sort: libstdc++-exact std::sort emulation on n keys
 * (ORBextractor.cc:674); perm[i] = input index that ends in slot i; force_depth >= 0 overrides the introsort depth limit
 * (exercises the heap fallback); on_device = 1 runs it in a CUDA thread as the quadtree kernel does.
 * orbx_debug_sincos: the kernel's sinf/cosf of angle_deg * (float)(pi/180) (ORBextractor.cc:86-88);
 * orbx_debug_atan2: the kernel's cv::fastAtan2 (ORBextractor.cc:76). */
int orbx_debug_sort(const uint32_t* keys, int n, int* perm, int force_depth, int on_device);
int orbx_debug_sincos(const float* angles_deg, float* out_sin, float* out_cos, int n);
int orbx_debug_atan2(const float* y, const float* x, float* out, int n);

const char* orbx_version(void);
const char* orbx_last_error(void); /* thread-local description of the last failure */

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
