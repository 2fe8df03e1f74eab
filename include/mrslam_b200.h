/* mrslam_b200.h -- C-ABI of the B200-native MR_SLAM hot path (libmrslam_b200.so).
 *
 * Plain C, opaque handles, plain pointers and sizes, int status codes (0 = OK, <0 = error; the
 * message of the last error on the calling thread is returned by mrslam_last_error()).  No
 * exceptions cross this boundary.  All 4x4 matrices are ROW-MAJOR float[16] / double[16]
 * (Eigen's default is column-major: the C++ adapters in include/ transpose).
 *
 * Every entry point names the reference interface it replaces (paths relative to the MR_SLAM
 * tree; FG = Mapping/src/fast_gicp, EM = Mapping/src/elevation_mapping_periodical/elevation_mapping,
 * BEV = LoopDetection/generate_bev_cython_binary, DISCO = LoopDetection/src/disco_ros,
 * TR = LoopDetection/torch-radon, RING = LoopDetection/src/RING_ros).
 *
 * There is NO CPU fallback behind this ABI: every compute entry point launches sm_100a kernels
 * and fails with MRSLAM_ERR_CUDA when no usable device is present.
 */
#ifndef MRSLAM_B200_H
#define MRSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRSLAM_OK 0
#define MRSLAM_ERR_CUDA (-1)     /* CUDA runtime error / no device */
#define MRSLAM_ERR_ARG (-2)      /* bad argument */
#define MRSLAM_ERR_STATE (-3)    /* call order (e.g. align before inputs are set) */
#define MRSLAM_ERR_NOTCONV (-4)  /* LM failed ("lm not converged!!" in the reference; result still readable) */

/* ------------------------------------------------------------------------------------------ */
/* runtime                                                                                     */
/* ------------------------------------------------------------------------------------------ */
int mrslam_init(int device);                 /* cudaSetDevice + stream/pool set-up; idempotent */
const char* mrslam_last_error(void);
const char* mrslam_version(void);
int mrslam_device_synchronize(void);
void* mrslam_stream(void);                   /* the cudaStream_t every kernel of this library is launched on */
/* instrumentation: kernels launched by this library since process start (bench.py "gpu_launches") */
unsigned long long mrslam_kernel_launch_count(void);
/* instrumentation: per-kernel device time with CUDA events on the library stream (ids: see ProfId in csrc/common.cuh) */
int mrslam_prof_enable(int on);
int mrslam_prof_get(int id, double* total_ms, long long* launches);

/* ------------------------------------------------------------------------------------------ */
/* FastGICP  --  replaces fast_gicp::FastGICP<PointSource,PointTarget>                          */
/*   FG/include/fast_gicp/gicp/fast_gicp.hpp:19-100, lsq_registration.hpp:15-85                 */
/*   FG/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:9-298, impl/lsq_registration_impl.hpp    */
/*   python: FG/src/python/main.cpp:152-190 (pygicp)                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct mrslam_gicp mrslam_gicp_t;
typedef struct mrslam_cloud mrslam_cloud_t;  /* device-resident cloud + spatial tiles + cached covariances */

/* RegularizationMethod, FG/include/fast_gicp/gicp/gicp_settings.hpp:6 (same order) */
enum { MRSLAM_REG_NONE = 0, MRSLAM_REG_MIN_EIG = 1, MRSLAM_REG_NORMALIZED_MIN_EIG = 2, MRSLAM_REG_PLANE = 3, MRSLAM_REG_FROBENIUS = 4 };
/* LSQ_OPTIMIZER_TYPE, FG/include/fast_gicp/gicp/lsq_registration.hpp:13 */
enum { MRSLAM_OPT_LM = 0, MRSLAM_OPT_GN = 1 };

/* clouds: `xyz` points at the first float of the first point, consecutive points are
 * `stride_floats` floats apart (3 = packed xyz, 4 = pcl::PointXYZ, 8 = pcl::PointXYZI / PointNormal). */
mrslam_cloud_t* mrslam_cloud_create_host(const float* xyz, int n, int stride_floats);
mrslam_cloud_t* mrslam_cloud_create_device(const float* d_xyz, int n, int stride_floats); /* device pointer (e.g. NCCL recv buffer) */
/* `count` clouds from device pointers packed by ONE kernel launch (a loop-closure batch out of an NCCL receive buffer);
 * out[count] receives the handles.  The inputs are consumed in stream order: they may be reused after the call returns
 * only by work ordered after mrslam_stream(). */
int mrslam_cloud_create_device_batch(const float* const* d_xyz, const int* n, int stride_floats, int count, mrslam_cloud_t** out);
void mrslam_cloud_destroy(mrslam_cloud_t*);
int mrslam_cloud_size(const mrslam_cloud_t*);

/* GlobalManager::mergeNearestKeyframes (Mapping/src/global_manager/src/global_manager.cpp:1894-1939), the front of ICPCheck:
 * `count` keyframe clouds (x,y,z[,intensity] points `stride_floats` apart; host pointers, or device pointers when on_device != 0)
 * are transformed by their row-major 4x4 T16[i] (= currPose^-1 * nearPose, computed by the caller), concatenated, cropped to
 * |x|,|y| <= xy_limit (pcl::PassThrough, 60 m in the node) and reduced by a pcl::VoxelGrid centroid filter with cubic `leaf`
 * (icp_filter_size).  Output: DEVICE buffer d_out_xyzi of `capacity` points (x,y,z,intensity), *n_out points written, voxels in
 * ascending voxel index like PCL; feed it to mrslam_cloud_create_device(d_out_xyzi, *n_out, 4).  capacity = sum n is always enough. */
int mrslam_keyframe_merge(const float* const* clouds, const int* n, int stride_floats, int count, const float* T16, float xy_limit,
                          float leaf, int on_device, float* d_out_xyzi, int capacity, int* n_out);

/* build_neighbors_NN(pointcloud, k) of the RING++ generator (LoopDetection/src/RING_ros/util.py:163-170, k = 30): exact kNN lists
 * (self first, (distance, index) order), sample covariances (row-major 3x3, divisor k-1), eigen-entropy and [e0 e1 e2 | f0 f1]
 * (3-D and x-y eigenvalues, descending) -- the inputs of voxelfeat.GPUFeatureExtractor.  HOST arrays; covs / entropy may be NULL. */
int mrslam_ringpp_build_neighbors(const float* xyz, int n, int stride_floats, int k, int* indices_out, float* covs_out9,
                                  float* entropy_out, float* eigens_out5);

/* Wire formats at the seams.  Keys of dislam_msgs/Loop {int64 id0, id1}: ((97 + robot) << 56) + keyframe index + 1
 * (global_manager.cpp:2587-2609 robotID2Key / Key2robotID, :1949-1951; RING/util.py:253-260). */
unsigned long long mrslam_key_pack(int robot, long long keyframe_index);
int mrslam_key_robot(unsigned long long key);
long long mrslam_key_index(unsigned long long key);
/* sensor_msgs/PointCloud2 data blob (dislam_msgs/SubMap.submap / .keyframePC; n_points = width*height records of point_step bytes,
 * float32 fields at byte offsets off_*; off_intensity < 0 = none) -> packed x,y,z,intensity in DEVICE memory (capacity n_points),
 * in record order, NaN records dropped when skip_nans (pc2.read_points(skip_nans=True), main_RING.py:251-254).  `data` is a host
 * pointer (on_device = 0) or a device pointer.  The result can go to mrslam_cloud_create_device(d_out_xyzi, *n_out, 4). */
int mrslam_pointcloud2_unpack(const unsigned char* data, int on_device, int n_points, int point_step, int off_x, int off_y, int off_z,
                              int off_intensity, int skip_nans, float* d_out_xyzi, int* n_out);

mrslam_gicp_t* mrslam_gicp_create(void);                              /* FastGICP() ctor: fast_gicp_impl.hpp:9-24 defaults */
void mrslam_gicp_destroy(mrslam_gicp_t*);
int mrslam_gicp_set_num_threads(mrslam_gicp_t*, int n);               /* setNumThreads: accepted, no effect on the GPU */
int mrslam_gicp_set_correspondence_randomness(mrslam_gicp_t*, int k); /* setCorrespondenceRandomness (1..32) */
int mrslam_gicp_set_regularization_method(mrslam_gicp_t*, int method);
int mrslam_gicp_set_max_correspondence_distance(mrslam_gicp_t*, double d); /* pcl::Registration::setMaxCorrespondenceDistance */
int mrslam_gicp_set_maximum_iterations(mrslam_gicp_t*, int n);        /* pcl::Registration::setMaximumIterations */
int mrslam_gicp_set_transformation_epsilon(mrslam_gicp_t*, double e); /* pcl::Registration::setTransformationEpsilon */
int mrslam_gicp_set_rotation_epsilon(mrslam_gicp_t*, double e);       /* LsqRegistration::setRotationEpsilon */
int mrslam_gicp_set_initial_lambda_factor(mrslam_gicp_t*, double f);  /* LsqRegistration::setInitialLambdaFactor */
int mrslam_gicp_set_lsq_optimizer(mrslam_gicp_t*, int type);
int mrslam_gicp_set_debug_print(mrslam_gicp_t*, int on);              /* LsqRegistration::setDebugPrint */

/* Host input: the upload is only ENQUEUED (copy stream) by these calls and is waited for / unpacked when the cloud is first used
 * (prepare / align).  From pageable memory -- what the reference's pcl clouds are -- the source buffer may be reused as soon as the
 * call returns; a PAGE-LOCKED source buffer is read asynchronously and must stay unchanged until that first use has completed.
 * Handing over the clouds of the next batch while the current one is being aligned is allowed and overlaps their uploads with it. */
int mrslam_gicp_set_input_source(mrslam_gicp_t*, const float* xyz, int n, int stride_floats); /* setInputSource (host) */
int mrslam_gicp_set_input_target(mrslam_gicp_t*, const float* xyz, int n, int stride_floats); /* setInputTarget (host) */
/* share a device-resident cloud (and its cached covariances) between registrations; not owned by the handle */
int mrslam_gicp_set_input_source_cloud(mrslam_gicp_t*, mrslam_cloud_t*);
int mrslam_gicp_set_input_target_cloud(mrslam_gicp_t*, mrslam_cloud_t*);
int mrslam_gicp_swap_source_and_target(mrslam_gicp_t*);               /* swapSourceAndTarget */
int mrslam_gicp_clear_source(mrslam_gicp_t*);                         /* clearSource */
int mrslam_gicp_clear_target(mrslam_gicp_t*);                         /* clearTarget */
/* set/getSource/TargetCovariances: n x 16 doubles, each a ROW-MAJOR 4x4 (4th row/col zero) */
int mrslam_gicp_set_source_covariances(mrslam_gicp_t*, const double* cov4x4, int n);
int mrslam_gicp_set_target_covariances(mrslam_gicp_t*, const double* cov4x4, int n);
int mrslam_gicp_get_source_covariances(mrslam_gicp_t*, double* cov4x4, int n); /* computes them if needed */
int mrslam_gicp_get_target_covariances(mrslam_gicp_t*, double* cov4x4, int n);

/* pcl::Registration::align(output, guess) -> FastGICP::computeTransformation (fast_gicp_impl.hpp:103-112).
 * guess16 may be NULL (identity).  Returns MRSLAM_OK or MRSLAM_ERR_NOTCONV. */
int mrslam_gicp_align(mrslam_gicp_t*, const float* guess16);
/* n registrations advanced in lock step (one kernel launch per LM phase for the whole batch):
 * the loop-closure-storm entry point (BASELINE configs 2 and 4).  guesses16: n x 16 or NULL. */
int mrslam_gicp_align_batch(mrslam_gicp_t** handles, int n, const float* guesses16);
/* Spatial structures + covariances (fast_gicp_impl.hpp:72-90 kd-tree build, :241-297 calculate_covariances) of the source
 * (which & 1) and / or target (which & 2) clouds of n registrations, ahead of align: lets a caller work on the clouds it
 * already holds while the others are still in flight on NVLink.  align / align_batch skip what is already prepared. */
int mrslam_gicp_prepare_batch(mrslam_gicp_t** handles, int n, int which);

int mrslam_gicp_get_final_transformation(const mrslam_gicp_t*, float* T16);   /* getFinalTransformation (x0.cast<float>()) */
int mrslam_gicp_get_final_transformation_d(const mrslam_gicp_t*, double* T16);
int mrslam_gicp_has_converged(const mrslam_gicp_t*);                          /* hasConverged */
int mrslam_gicp_get_num_iterations(const mrslam_gicp_t*);                     /* nr_iterations_ */
int mrslam_gicp_get_final_hessian(const mrslam_gicp_t*, double* H36);         /* getFinalHessian */
int mrslam_gicp_get_eval_counts(const mrslam_gicp_t*, int* n_linearize, int* n_error);
/* pcl::Registration::getFitnessScore(max_range): mean NN squared distance <= max_range after final_transformation_ */
double mrslam_gicp_get_fitness_score(mrslam_gicp_t*, double max_range);
int mrslam_gicp_get_fitness_score_batch(mrslam_gicp_t** handles, int n, double max_range, double* out);
/* `output` cloud of align(): the source transformed by final_transformation_ (lsq_registration_impl.hpp:78) */
int mrslam_gicp_get_aligned(mrslam_gicp_t*, float* xyz_out, int stride_floats);
/* LsqRegistration::evaluateCost(relative_pose, H, b) -> linearize (H36/b6 may be NULL) */
int mrslam_gicp_evaluate_cost(mrslam_gicp_t*, const float* T16, double* H36, double* b6, double* cost);
/* test hooks: state of the last linearisation in SOURCE ORIGINAL ORDER (corr = target original index or -1) */
int mrslam_gicp_debug_last_correspondences(mrslam_gicp_t*, int* corr, float* sq_dist, double* mahalanobis9);
int mrslam_gicp_compute_error(mrslam_gicp_t*, const double* T16, double* cost); /* FastGICP::compute_error */
/* instrumentation: distance evaluations (0: kNN, 1: NN-1) and k-min insertion rounds (2) since stats_enable(1) */
int mrslam_gicp_stats_enable(int on);
/* development aid: with MRSLAM_BUILD_DIAG=1 in the environment, print (stderr) and reset the per-stage device times of the cloud build */
void mrslam_build_diag_dump(const char* label);
unsigned long long mrslam_gicp_stats_get(int which);
unsigned long long mrslam_gicp_stats_pairs(void);
/* kNN building block (calculate_covariances' search): indices (original order) and squared distances, n x k */
int mrslam_cloud_knn(mrslam_cloud_t*, int k, int* idx_out, float* sqd_out);

/* ------------------------------------------------------------------------------------------ */
/* Elevation mapping -- replaces the free functions of EM/cuda/gpu_process.cu (libgpu.so)      */
/*   declared at EM/src/ElevationMapping.cpp:44-50, SensorProcessorBase.cpp:34,                 */
/*   RobotMotionMapUpdater.cpp:18.  One hidden map per handle (the reference: one per process). */
/* ------------------------------------------------------------------------------------------ */
typedef struct mrslam_elev mrslam_elev_t;
mrslam_elev_t* mrslam_elev_create(void);
void mrslam_elev_destroy(mrslam_elev_t*);
/* Init_GPU_elevationmap(length, resolution, mahalanobisThr, obstacle_threshold)  gpu_process.cu:938 */
int mrslam_elev_init(mrslam_elev_t*, int length, float resolution, float mahalanobis_thr, float obstacle_thr);
/* Move(current_Position[3], resolution, length, central_coordinate_out[2], start_indice_out[2], position_shift_out[2])  :1002 */
int mrslam_elev_move(mrslam_elev_t*, const float* current_position3, float resolution, int length, float* central_out2,
                     int* start_out2, float* shift_out2);
/* Process_points(map_index, x, y, z, var, x_ts, y_ts, z_ts, T, n, lower, upper, min_r, beam_a, beam_c,
 *                sensorJacobian, rotationVariance, C_SB_transpose, P_mul_C_BM_transpose, B_r_BS_skew)   :1083
 * Eigen arguments are passed as ROW-MAJOR float arrays (T 4x4, 1x3, 3x3, 3x3, 1x3, 3x3). All point arrays are HOST. */
int mrslam_elev_process_points(mrslam_elev_t*, int* map_index, const float* x, const float* y, const float* z, float* var,
                               float* x_ts, float* y_ts, float* z_ts, const float* T16, int n, double lower, double upper,
                               float min_r, float beam_a, float beam_c, const float* sensorJacobian3,
                               const float* rotationVariance9, const float* C_SB_transpose9,
                               const float* P_mul_C_BM_transpose3, const float* B_r_BS_skew9);
/* Fuse(length, n, idx, R, G, B, intensity, height, var)   :1152   (host arrays); returns #cell fuses in *n_fused (may be NULL) */
int mrslam_elev_fuse(mrslam_elev_t*, int length, int n, const int* idx, const int* R, const int* G, const int* B,
                     const float* intensity, const float* height, const float* var, long long* n_fused);
/* device-resident fast path: Process_points + Fuse in one call on device arrays (x,y,z SoA device pointers;
 * colour/intensity may be NULL = 0).  near_filter=1 keeps the reference's hard-coded sensor-frame filter (:393). */
int mrslam_elev_process_fuse_device(mrslam_elev_t*, const float* d_x, const float* d_y, const float* d_z, int n,
                                    const float* T16, double lower, double upper, float min_r, float beam_a, float beam_c,
                                    const float* sensorJacobian3, const float* rotationVariance9, const float* C_SB_transpose9,
                                    const float* P_mul_C_BM_transpose3, const float* B_r_BS_skew9, int near_filter,
                                    long long* n_fused);
/* One frame of the batched entry point below: Move(position) when do_move != 0, then Process_points + Fuse of the device
 * arrays with the same arguments as the per-frame call. */
typedef struct mrslam_elev_frame {
  const float *d_x, *d_y, *d_z;
  int n;
  int do_move;
  float position[3];
  float T16[16];
  double lower, upper;
  float min_r, beam_a, beam_c;
  float sJ[3], rotVar[9], CSBT[9], PCBM[3], BrBS[9];
  int near_filter;
} mrslam_elev_frame_t;
/* nframes frames in call order (map replay, or the frames queued while a loop closure was processed): same results as
 * nframes x (Move, process_fuse_device), bit for bit, but fused in ONE pass per group of <= 64 frames: the Moves are planned on
 * the host, all frames are processed by one launch, all their points are sorted by cell once (library's own stable radix sort,
 * invalid points dropped in its first pass), and one run walk applies the per-frame clamps and the Moves' clears in order.
 * No host synchronisation; counts go to fused_total. */
int mrslam_elev_process_fuse_device_batch(mrslam_elev_t*, const mrslam_elev_frame_t* frames, int nframes, float resolution, int length);
/* The same replay with its map-writing half on the map's own stream instead of mrslam_stream(): the call returns after the
 * enqueue and an unrelated batch (a loop-closure storm on the library stream) runs concurrently with it.  Inputs must be
 * ready in mrslam_stream() order at the call; every later mrslam_elev_* call on this map joins the replay first. */
int mrslam_elev_process_fuse_device_batch_async(mrslam_elev_t*, const mrslam_elev_frame_t* frames, int nframes, float resolution,
                                                int length);
/* cell fuses accumulated on the device by process_fuse_device calls that passed n_fused == NULL (no per-frame host sync) */
int mrslam_elev_fused_total(mrslam_elev_t*, int reset, long long* total);
int mrslam_elev_mapvar_update(mrslam_elev_t*, int length, float var_update);          /* Mapvar_update  :1144 */
int mrslam_elev_map_optmove(mrslam_elev_t*, const float* opt_p, float height_update, float resolution, int length,
                            float* opt_aligned_position2);                             /* Map_optmove    :1213 */
int mrslam_elev_map_closeloop(mrslam_elev_t*, const float* update_position, float height_update, int length,
                              float resolution);                                       /* Map_closeloop  :1233 */
/* Map_feature(length, elevation, var, R, G, B, rough, slope, traver, intensity)   :1254  (host arrays, any may be NULL):
 * 5x5 plane-fit traversability (G_Mapfeature :547-668) + read-out of the map; cells without elevation get 0 / 0 / -10. */
int mrslam_elev_map_feature(mrslam_elev_t*, int length, float* elevation, float* var, int* colorR, int* colorG, int* colorB,
                            float* rough, float* slope, float* traver, float* intensity);
/* Tile-sharded global grid (BASELINE config 5): restrict this instance to the STORAGE rows [row_lo, row_hi): points
 * falling into other rows are treated as outside the map by Process_points / process_fuse_device (index -1).  Cells are disjoint
 * across tiles, so the union of the tiles equals the single-GPU map bit for bit; no reduction is needed for the fuse. */
int mrslam_elev_set_owned_rows(mrslam_elev_t*, int row_lo, int row_hi);
/* n_rows consecutive storage rows of one plane <-> DEVICE buffer (n_rows*length 4-byte elements): halo rows of the neighbouring
 * tiles for Map_feature's 5x5 window, or whole tile blocks when the full map is assembled.  plane: 0 elevation, 1 variance,
 * 2 intensity, 3 lowest, 4 traversability, 5/6/7 R/G/B.  to_map != 0: buffer -> map.  Ordered on mrslam_stream(). */
int mrslam_elev_rows_copy(mrslam_elev_t*, int plane, int storage_row, int n_rows, void* d_buf, int to_map);
/* Raytracing(length)   :1301   obstacle clean-up along the sensor rays (G_Raytracing :706-889) + reset of map_lowest */
int mrslam_elev_raytracing(mrslam_elev_t*, int length);
/* Stable LSD radix sort of (u32 key, u32 value) pairs by key bits [0, bits) with the library's own kernels (the sort behind the
 * elevation replay, csrc/radix_sort.cuh; replaces cub::DeviceRadixSort there): keys equal to drop_key are removed when use_drop != 0
 * (how invalid points leave the fuse).  Host arrays in and out; *n_out = number of pairs kept.  Utility / test entry point. */
int mrslam_sort_pairs_u32(const unsigned* keys, const unsigned* vals, int n, int bits, unsigned drop_key, int use_drop,
                          unsigned* keys_out, unsigned* vals_out, int* n_out);
/* read back the hidden map (any pointer may be NULL); length*length entries each, storage (ring-buffer) order */
int mrslam_elev_get_map(mrslam_elev_t*, float* elevation, float* variance, float* intensity, int* R, int* G, int* B,
                        float* lowest);
int mrslam_elev_get_state(mrslam_elev_t*, float* central2, int* start2);

/* ------------------------------------------------------------------------------------------ */
/* BEV voxelisers -- replace GPUTransformer of voxelocc (BEV/src/manager.hh, wrapper.pyx:13-39) */
/* and gputransform (DISCO/tools/multi-layer-polar-gpu/cython/src/manager.hh).                  */
/* point: SoA float[3*size] (x.., y.., z..) HOST; out: HOST float[num_x*num_y*num_height*enough_large*3] */
/* ------------------------------------------------------------------------------------------ */
int mrslam_voxelocc_transform(const float* point, int size, int max_length, int max_height, int num_x, int num_y,
                              int num_height, int enough_large, float* out);
int mrslam_polar_transform(const float* point, int size, int max_length, int max_height, int num_ring, int num_sector,
                           int num_height, int enough_large, float* out);
/* voxelfeat (RING++) -- BEVF = LoopDetection/generate_bev_pointfeat_cython: GPUFeatureExtractor::get_features
 * (BEVF/src/manager.cu:127-181, kernel.cu:16-104; point AoS xyz, neighbours int[size*k], eigens float[size*5], out float[size*featsize])
 * and the feature-BEV GPUTransformer (BEVF/src/manager.cu:18-101, kernel.cu:106-164; point SoA with featsize channels,
 * out float[num_x*num_y*featsize]; num_height must be 1).  Host pointers. */
int mrslam_voxelfeat_features(const float* point_aos, int size, int featsize, int k, const int* neighbors, const float* eigens,
                              float* feature_out);
int mrslam_voxelfeat_transform(const float* point, int size, int max_length, int max_height, int num_x, int num_y,
                               int num_height, int featsize, float* out);
/* device batch variant used by the RING pipeline: points SoA per scan on the device, offsets[b]..offsets[b+1];
 * writes only what generate_RING keeps (channel 2, the max-height image) as d_bev[batch][num_x][num_y]. */
int mrslam_voxelocc_bev_device(const float* d_points_soa, const long long* offsets_host, int batch, int max_length,
                               int max_height, int num_x, int num_y, float* d_bev);
/* RING front end from RAW scans in one pass: load_pc_infer (LoopDetection/src/RING_ros/util.py:91-112: keep |x|,|y| < xy_range,
 * 0 < z < z_range, scale to [-1,1]) fused with the channel-2 (max-height) voxeliser of generate_RING (util.py:174-188).
 * d_points: device, `stride` floats per point (x y z first), scans concatenated; offsets_host[batch+1] in points.  d_bev:
 * device float[batch*num_x*num_y].  The reference does both steps on the CPU per scan. */
int mrslam_ring_scans_to_bev_device(const float* d_points, const long long* offsets_host, int batch, int stride, float xy_range,
                                    float z_range, int max_length, int max_height, int num_x, int num_y, float* d_bev);

/* ------------------------------------------------------------------------------------------ */
/* Radon / RING -- replace torch_radon.ParallelBeam.forward (TR/torch_radon/radon.py:62-87,      */
/* TR/src/forward.cu:12-178) and the RING descriptor tail + matcher (RING/util.py:192-198,362-374) */
/* All pointers are DEVICE pointers (torch CUDA tensors' data_ptr on the python side).            */
/* ------------------------------------------------------------------------------------------ */
/* interp: 0 = exact fp32 bilinear, 1 = texture-unit emulation (8-bit fractional weights);
 * fast_trig: 1 = __sinf/__cosf as the reference kernel, 0 = sinf/cosf. */
int mrslam_radon_forward(const float* d_img, int batch, int height, int width, const float* d_angles, int n_angles,
                         int det_count, float det_spacing, int interp, int fast_trig, float* d_sino);
/* normalise (mean / unbiased std over each [C,A,D] item) + ortho FFT along the angle axis; out complex64 interleaved */
int mrslam_ring_tiring(const float* d_sino, int items, int channels, int n_angles, int det_count, float* d_tiring);
/* fast_corr of query q against `n_db` database descriptors: dist[n_db], angle[n_db] (device) */
int mrslam_ring_fast_corr(const float* d_query, const float* d_db, int n_db, int channels, int n_angles, int det_count,
                          float* d_dist, int* d_angle, float* d_corr_opt);
/* solve_translation(query, positive, rot_angle) (RING/util.py:388-423) with the row roll of main_RING.py:171-172 folded in:
 * per sinogram row the circular cross-correlation shift along the detector axis (b[n_angles], optional device output), then the
 * least-squares x, y and the residual norm -> xy_err3 (HOST, 3 floats).  d_query / d_positive: float [channels, n_angles, det]. */
int mrslam_ring_solve_translation(const float* d_query, const float* d_positive, int channels, int n_angles, int det_count,
                                  float rot_angle, int row_shift, float* d_row_shift_opt, float* xy_err3);
/* GlobalManager::calcRelOri(newDiSCO, oldDiSCO) (Mapping/src/global_manager/src/global_manager.cpp:2719-2762) for one new
 * signature a against n_b old ones b: real/imag planes float [height, width] (a) and [n_b, height, width] (b), DEVICE pointers.
 * d_rel_angle[n_b] = width offset of the arg-max of Re(IDFT2(cross spectrum)) * 3 degrees; d_max_loc_opt[n_b] = the flat arg-max. */
int mrslam_disco_rel_ori(const float* d_real_a, const float* d_imag_a, const float* d_real_b, const float* d_imag_b, int n_b,
                         int height, int width, float* d_rel_angle, int* d_max_loc_opt);
/* fused generate_RING for a batch of BEV images: Radon -> normalise -> FFT (one kernel per image) */
int mrslam_ring_generate(const float* d_bev, int batch, int size, const float* d_angles, int n_angles, int interp,
                         int fast_trig, float* d_sino_opt, float* d_tiring);

#ifdef __cplusplus
}
#endif
#endif /* MRSLAM_B200_H */
