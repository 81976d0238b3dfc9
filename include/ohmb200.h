/*
 * ohmb200.h — C ABI of libohmb200.so: B200-native (sm_100a) batched ray integration for ohm maps.
 *
 * This is the drop-in boundary for ohm's GPU ray-integration path.  Each entry point cites the
 * reference interface it replaces (paths relative to csiro-robotics/ohm @ 4e2e769).  The C++ facade in
 * include/ohmb200/GpuMap.hpp (ohm::RayMapper / GpuMap / GpuNdtMap / GpuTsdfMap signatures) is a thin
 * wrapper over these calls; INTEGRATION.md shows the binding a maintainer adds behind OhmAppGpu.
 *
 * Conventions: plain pointers and sizes only; `int` returns 0 on success and a negative OHMB200_E_* code on
 * failure (message via ohmb200_last_error()); size_t returns follow the reference call they replace.
 * Not thread-safe per map (same as ohmgpu/GpuMap.h:116-125): one caller thread per map.
 */
#ifndef OHMB200_H
#define OHMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define OHMB200_API
#else
#define OHMB200_API __attribute__((visibility("default")))
#endif

/* Voxel layers — ohm/DefaultLayer.cpp:76-337.  Bit (1u << id) in ohmb200_params.layers enables a layer. */
enum ohmb200_layer
{
  OHMB200_LAYER_OCCUPANCY = 0,  /* f32, clear value +inf             (DefaultLayer.cpp:85-91)   */
  OHMB200_LAYER_MEAN = 1,       /* {u32 coord, u32 count}            (VoxelMeanCompute.h:29-33) */
  OHMB200_LAYER_TRAVERSAL = 2,  /* f32                               (DefaultLayer.cpp:132)     */
  OHMB200_LAYER_TOUCH_TIME = 3, /* u32 ms since first ray            (VoxelTouchTimeCompute.h)  */
  OHMB200_LAYER_INCIDENT = 4,   /* u32 packed normal                 (VoxelIncidentCompute.h)   */
  OHMB200_LAYER_COVARIANCE = 5, /* 6 x f32 packed sqrt covariance    (CovarianceVoxelCompute.h:56-64) */
  OHMB200_LAYER_INTENSITY = 6,  /* {f32 mean, f32 cov}               (CovarianceVoxelCompute.h:67-73) */
  OHMB200_LAYER_HIT_MISS = 7,   /* {u32 hit, u32 miss}               (CovarianceVoxelCompute.h:76-82) */
  OHMB200_LAYER_TSDF = 8,       /* {f32 weight, f32 distance}        (VoxelTsdfCompute.h:20-24) */
  OHMB200_LAYER_SECONDARY = 9,  /* {f32 m2, u16 range_mean, u16 count} (VoxelSecondarySample.h:29-38) */
  OHMB200_LAYER_COUNT = 10
};

/* Which mapper the map runs — the classes OhmAppGpu::prepareForRun picks between (ohmapp/OhmAppGpu.cpp:164-259). */
enum ohmb200_mode
{
  OHMB200_MODE_OCCUPANCY = 0, /* ohm::GpuMap      (ohmgpu/GpuMap.h:143)                         */
  OHMB200_MODE_NDT = 1,       /* ohm::GpuNdtMap, NdtMode::kOccupancy (ohmgpu/GpuNdtMap.h:63)    */
  OHMB200_MODE_NDT_TM = 2,    /* ohm::GpuNdtMap, NdtMode::kTraversability                       */
  OHMB200_MODE_TSDF = 3       /* ohm::GpuTsdfMap  (ohmgpu/GpuTsdfMap.h:37)                      */
};

/* Ray filters — ohm/RayFilter.cpp:15-55 (the std::function the map holds, OccupancyMap.cpp:215-218). */
enum ohmb200_filter
{
  OHMB200_FILTER_NONE = 0,
  OHMB200_FILTER_GOOD_RAY = 1,  /* goodRayFilter(max_range); the OccupancyMap default, range 1e10 */
  OHMB200_FILTER_CLIP_RANGE = 2, /* clipRayFilter(max_length): clip + kRffClippedEnd              */
  OHMB200_FILTER_CLIP_BOX = 3    /* clipBounded(Aabb clip_box) (ohm/RayFilter.cpp:57-76)           */
};

/* RayFlag — ohm/RayFlag.h:16-60 (same bit values). */
enum ohmb200_ray_flag
{
  OHMB200_RF_DEFAULT = 0,
  OHMB200_RF_END_POINT_AS_FREE = 1u << 0,
  OHMB200_RF_STOP_ON_FIRST_OCCUPIED = 1u << 1, /* order-dependent across rays: integrated by one thread, ray after ray,
                                                * exactly as RayMapperOccupancy.cpp:183,234 (ClearingPattern-sized batches);
                                                * refused (OHMB200_E_INVALID) on a sharded or paged-out map */
  OHMB200_RF_EXCLUDE_ORIGIN = 1u << 2,
  OHMB200_RF_EXCLUDE_SAMPLE = 1u << 3,
  OHMB200_RF_EXCLUDE_RAY = 1u << 4,
  OHMB200_RF_EXCLUDE_UNOBSERVED = 1u << 5,
  OHMB200_RF_EXCLUDE_FREE = 1u << 6,
  OHMB200_RF_EXCLUDE_OCCUPIED = 1u << 7,
  OHMB200_RF_REVERSE_WALK = 1u << 8 /* accepted and ignored, as the CPU mappers do */
};

enum ohmb200_error
{
  OHMB200_OK = 0,
  OHMB200_E_INVALID = -1,     /* bad argument */
  OHMB200_E_CUDA = -2,        /* CUDA runtime failure (gputil::ApiException in the reference) */
  OHMB200_E_NO_DEVICE = -3,   /* no sm_100 device: gpuOk() == false, calls are no-ops (GpuMap.cpp:548-551) */
  OHMB200_E_CACHE_FULL = -4,  /* region table full (GpuLayerCache kCacheFull, GpuMap.cpp:935-975) */
  OHMB200_E_NOT_FOUND = -5,   /* region/layer absent */
  OHMB200_E_OVERFLOW = -6     /* reported once by ohmb200_sync.  "segment list": a batch cut into more region segments than
                               * its list holds (tiny regions, very long rays) was dropped WHOLE - nothing of it was applied -
                               * and the list has been doubled: integrate that batch again.  "ordered records": too many
                               * misses on voxels that are also hit in the same batch; that batch's results are incomplete,
                               * the list has been doubled for the batches that follow */
};

/* Map + mapper parameters.  Mirrors OccupancyMapDetail (ohm/private/OccupancyMapDetail.h:47-84),
 * NdtMapDetail (ohm/private/NdtMapDetail.h:24-40) and TsdfOptions (ohm/VoxelTsdf.h:27-37). */
typedef struct ohmb200_params
{
  double resolution;        /* voxel edge, metres */
  int32_t region_dim[3];    /* voxels per region edge; 32 by default (OccupancyMap.h:24-26) */
  double origin[3];         /* map origin */
  float hit_value;          /* log-odds added on a hit  (logit 0.9)  */
  float miss_value;         /* log-odds added on a miss (logit 0.45) */
  float min_value;          /* -2.0   */
  float max_value;          /* 3.511  */
  float threshold_value;    /* occupancy threshold (logit 0.5 = 0) */
  int32_t saturate_min;     /* saturateAtMinValue */
  int32_t saturate_max;     /* saturateAtMaxValue */
  uint32_t layers;          /* bitset of (1u << OHMB200_LAYER_*) */
  int32_t filter_kind;      /* ohmb200_filter */
  double filter_range;      /* max_range / max_length for the filter */
  double clip_box[6];       /* OHMB200_FILTER_CLIP_BOX: min xyz, max xyz */
  float sensor_noise;       /* NDT */
  float adaptation_rate;
  float reinit_threshold;
  uint32_t reinit_count;
  uint32_t sample_threshold;
  float initial_intensity_cov;
  int32_t ndt_tm;
  float tsdf_max_weight;    /* TSDF */
  float tsdf_trunc;
  float tsdf_dropoff;
  float tsdf_sparsity;
} ohmb200_params;

/* Counters of the last integrate call and totals (the units of the roofline byte model). */
typedef struct ohmb200_stats
{
  uint64_t rays_in;          /* rays offered over the map's life */
  uint64_t rays_accepted;    /* rays that passed the filter */
  uint64_t voxel_visits;     /* V: DDA voxel visits (miss side; TSDF: all) */
  uint64_t sample_updates;   /* S: sample voxel updates */
  uint64_t ordered_records;  /* visits replayed in exact ray order because the voxel was also hit in the batch */
  uint64_t regions;          /* regions resident on the device */
  uint64_t region_capacity;  /* region slots allocated */
  uint64_t batches;          /* integrate calls */
  uint64_t kernel_launches;  /* kernels launched by this library */
  uint64_t sample_voxels;    /* S': distinct sample voxels per batch, summed over batches (NDT byte model) */
  uint64_t owned_visits;     /* voxel visits applied to THIS map's regions (== voxel_visits unless the exchange routes
                              * other ranks' segments here: then voxel_visits counts what this rank's own rays cut) */
} ohmb200_stats;

typedef struct ohmb200_map ohmb200_map;

/* Number of usable (compute capability 10.x) devices.  Replaces ohm::configureGpuFromArgs / gpuDevice()
 * (ohmgpu/OhmGpu.h:40). */
OHMB200_API int ohmb200_device_count(void);

/* Defaults of ohm::OccupancyMap's constructor (ohm/OccupancyMap.cpp:195-222), NdtMap and TsdfOptions. */
OHMB200_API void ohmb200_default_params(ohmb200_params *params, double resolution);

/* Construct map + mapper on `device`.  Replaces `new ohm::OccupancyMap(res, region_dim, flags)` followed by
 * `ohm::GpuMap(map, borrowed, expected_element_count, gpu_mem_size)` / GpuNdtMap / GpuTsdfMap
 * (ohmgpu/GpuMap.h:167-169, GpuNdtMap.h:72-74, GpuTsdfMap.h:46-48).  `device_bytes` is the GPU cache budget
 * (gpu_mem_size; 0 = default), which fixes the number of resident region slots.  NULL on failure. */
OHMB200_API ohmb200_map *ohmb200_create(const ohmb200_params *params, int mode, size_t device_bytes, int device);

/* ~GpuMap + ~OccupancyMap. */
OHMB200_API void ohmb200_destroy(ohmb200_map *map);

/* Update the mutable parameters (hit/miss/min/max/threshold/saturation, filter, NDT, TSDF).  Replaces
 * OccupancyMap::setHitValue/setMissValue/... , GpuMap::setRayFilter (GpuMap.h:214), GpuNdtMap::setSensorNoise
 * (GpuNdtMap.h:91), GpuTsdfMap::setTsdfOptions (GpuTsdfMap.h:64).  Geometry and layers must not change. */
OHMB200_API int ohmb200_set_params(ohmb200_map *map, const ohmb200_params *params);
OHMB200_API int ohmb200_get_params(const ohmb200_map *map, ohmb200_params *params);

/* ohm::RayMapper::integrateRays (ohm/RayMapper.h:57-58) as implemented by GpuMap::integrateRays
 * (ohmgpu/GpuMap.cpp:540-875).  `rays` = interleaved (origin, sample) f64 xyz triples in HOST memory,
 * `element_count` = 2 x ray count; `intensities` / `timestamps` are per ray and nullable.  Asynchronous with
 * respect to the device: inputs are copied to device staging before return.  Returns element_count (what the CPU
 * mappers return, RayMapperOccupancy.cpp:338) or 0 on failure; the reference's GpuMap returns 2 x the rays its host-side
 * filter accepted (GpuMap.cpp:874) - here the filter runs on the device after the call has returned, and callers only
 * test for zero; the accepted count is ohmb200_stats.rays_accepted. */
OHMB200_API size_t ohmb200_integrate(ohmb200_map *map, const double *rays, size_t element_count,
                                     const float *intensities, const double *timestamps, unsigned ray_flags);

/* Same call with the arrays already resident in device memory (HBM) on the map's device. */
OHMB200_API size_t ohmb200_integrate_device(ohmb200_map *map, const double *d_rays, size_t element_count,
                                            const float *d_intensities, const double *d_timestamps,
                                            unsigned ray_flags);

/* GpuMap::syncVoxels (ohmgpu/GpuMap.h:199): block until all queued device work is complete.  Voxel data stays
 * resident; read it back with ohmb200_read_region(s). */
OHMB200_API int ohmb200_sync(ohmb200_map *map);

/* OccupancyMap::regionCount / region enumeration.  Keys are written sorted by (z, y, x). */
OHMB200_API size_t ohmb200_region_count(ohmb200_map *map);
OHMB200_API size_t ohmb200_enumerate_regions(ohmb200_map *map, int16_t *keys_xyz, size_t capacity);

/* Bytes of one region chunk of `layer` (MapLayer::layerByteSize). */
OHMB200_API size_t ohmb200_region_layer_bytes(const ohmb200_map *map, int layer);

/* GpuLayerCache::syncToExternal(dst, dst_size, region_key) (ohmgpu/GpuLayerCache.h:291): copy one region chunk
 * of one layer to host memory (x + y*dx + z*dx*dy voxel order, ohm/MapChunk.h:47-50). */
OHMB200_API int ohmb200_read_region(ohmb200_map *map, const int16_t key_xyz[3], int layer, void *dst, size_t bytes);

/* GpuLayerCache::syncToMainMemory() (ohmgpu/GpuLayerCache.h:268) for `count` regions in one gather + one D2H:
 * dst receives count x region_layer_bytes, in the order of `keys_xyz`. */
OHMB200_API int ohmb200_read_regions(ohmb200_map *map, int layer, const int16_t *keys_xyz, size_t count, void *dst,
                                     size_t bytes);

/* The asynchronous form of the same (GpuLayerCache::syncToMainMemory queues the downloads and returns; the wait is
 * GpuLayerCache::updateEvents / GpuMap::syncVoxels, ohmgpu/GpuLayerCache.cpp:229-241,300-321,670-713): the chunks are
 * snapshotted on the device in stream order — batches integrated after this call do not show in them — and copied to
 * `dst` on a separate stream while those batches run.  `dst` (pinned memory for a real overlap) must stay valid until
 * ohmb200_download_wait returns; at most two downloads are in flight, a third call waits for the first.
 * A key that is not resident is reported by ohmb200_download_wait (OHMB200_E_NOT_FOUND; its chunk is left untouched). */
OHMB200_API int ohmb200_read_regions_async(ohmb200_map *map, int layer, const int16_t *keys_xyz, size_t count,
                                           void *dst, size_t bytes);
OHMB200_API int ohmb200_download_wait(ohmb200_map *map);

/* ohm::RayMapperSecondarySample::integrateRays (ohm/RayMapperSecondarySample.cpp:37-74), the dual-return mapper ohmpop
 * runs beside the main one: for every ray [primary sample, secondary sample] a Welford update (addSecondarySample,
 * ohm/VoxelSecondarySample.h:87-99) of the voxel holding the SECOND point with range = |second - first|; samples of one
 * voxel are applied in ray order.  No ray filter, no walk.  The map must have been created with
 * OHMB200_LAYER_SECONDARY in params.layers (MapFlag::kSecondarySample).  Returns the number of elements consumed. */
OHMB200_API size_t ohmb200_integrate_secondary(ohmb200_map *map, const double *rays, size_t element_count);
OHMB200_API size_t ohmb200_integrate_secondary_device(ohmb200_map *map, const double *d_rays, size_t element_count);

/* ohm::RaysQuery / ohm::RaysQueryGpu (ohm/RaysQuery.h:42-139, ohmgpu/RaysQueryGpu.h, kernel ohmgpu/gpu/RaysQuery.cl): for
 * each ray [origin, end] walk the resident map until the first occupied voxel (value > threshold).  Outputs per ray:
 * ranges = exit range of the last voxel that is not occupied (a float, widened — Query::ranges() is double),
 * unobserved_volumes = volume_coefficient * sum over unobserved voxels of (exit^3 - enter^3) (RaysQuery.h:23-40),
 * terminal_states = ohm::OccupancyType of the last voxel looked at (-2 null, -1 unobserved, 0 free, 1 occupied),
 * terminal_keys = its key {region x,y,z, local x,y,z}.  Rays the map's filter rejects report 0, 0, null, zeros.
 * Runs in stream order after every batch queued before it.  The _device form takes and fills device memory and
 * returns after queueing. */
OHMB200_API int ohmb200_rays_query(ohmb200_map *map, const double *rays, size_t element_count,
                                   double volume_coefficient, double *ranges, double *unobserved_volumes,
                                   int *terminal_states, int32_t *terminal_keys);
OHMB200_API int ohmb200_rays_query_device(ohmb200_map *map, const double *d_rays, size_t element_count,
                                          double volume_coefficient, double *d_ranges, double *d_unobserved_volumes,
                                          int *d_terminal_states, int32_t *d_terminal_keys);

/* ohm::LineKeysQuery / ohm::LineKeysQueryGpu (ohm/LineKeysQuery.h:20-90, ohmgpu/LineKeysQueryGpu.h, kernel
 * ohmgpu/gpu/LineKeys.cl): the voxel keys along each line [start, end], both end voxels included
 * (calculateSegmentKeys, ohm/CalculateSegmentKeys.cpp).  For line i the keys are keys[6 * result_indices[i] ...], six
 * int32 each {region x,y,z, local x,y,z}, result_counts[i] of them.  *total_keys receives the sum of the counts; keys
 * are written only when key_capacity (in keys) >= *total_keys — call once with keys = NULL, key_capacity = 0 to size the
 * buffer.  Uses only the map's geometry (resolution, region dimensions, origin). */
OHMB200_API int ohmb200_line_keys_query(ohmb200_map *map, const double *rays, size_t element_count,
                                        uint64_t *result_indices, uint64_t *result_counts, int32_t *keys,
                                        size_t key_capacity, size_t *total_keys);

/* GpuLayerCache::upload (ohmgpu/GpuLayerCache.cpp:172-182): make the region resident (creating it if absent)
 * and overwrite one layer chunk from host memory. */
OHMB200_API int ohmb200_write_region(ohmb200_map *map, const int16_t key_xyz[3], int layer, const void *src,
                                     size_t bytes);

/* GpuCache::clear (ohmgpu/GpuCache.h:109) + OccupancyMap::clear: drop every region. */
OHMB200_API int ohmb200_clear(ohmb200_map *map);

/* OccupancyMap::firstRayTime / setFirstRayTime (ohm/OccupancyMap.cpp:333-347). < 0 = unset. */
OHMB200_API double ohmb200_first_ray_time(const ohmb200_map *map);
OHMB200_API int ohmb200_set_first_ray_time(ohmb200_map *map, double t);

OHMB200_API int ohmb200_get_stats(ohmb200_map *map, ohmb200_stats *stats);

/* Use an external CUDA stream (a cudaStream_t passed as void*) instead of the map's own.  NULL restores it. */
OHMB200_API int ohmb200_set_stream(ohmb200_map *map, void *cuda_stream);

/* Multi-GPU sharding (new capability; the reference is single-device).  The map keeps only the regions
 * whose owner is `rank` of `world`: owner = (rx + 2 ry + 4 rz) mod world.  Every GPU is handed the
 * same rays (one NCCL all-gather per batch, done by the caller) and applies exactly the visits and samples that
 * fall in its own regions, so the union of the per-GPU maps is bit-identical to the single-GPU map. */
OHMB200_API int ohmb200_set_partition(ohmb200_map *map, int rank, int world);
OHMB200_API int ohmb200_region_owner(const int16_t key_xyz[3], int world);

/* ---- multi-GPU: the routed exchange (new capability; SURVEY §8e second option: "all-to-all of pre-cut segments") --------
 * ohmb200_set_partition above shards the MAP but hands every GPU every ray.  The exchange shards the WORK too: each rank
 * filters and cuts only ITS OWN rays and stores every region segment and every sample straight into the inbox of the GPU
 * that owns the region (peer stores over NVLink into one device allocation per rank, exported as a CUDA IPC handle; the
 * per-ray walk constants follow by copy engine).  The owner bins what it received and runs the single-GPU walk and
 * replay kernels.  The cut is the exact region-boundary cut of the single-GPU path — the exact counterpart of the
 * reference's clipped-end ray segments (ohmgpu/GpuMap.cpp:747-795, ohmgpu/gpu/AdjustOccupancy.cl:13-18) — so the union
 * of the per-GPU maps equals one GPU, or the CPU mapper, integrating rank 0's rays, then rank 1's, ... bit for bit.
 *
 *   every rank:  ohmb200_exchange_open(map, rank, world, max_rays_per_rank, &mine)
 *                all-gather the handles (any transport: torch.distributed, MPI, a pipe), in rank order
 *                ohmb200_exchange_connect(map, handles, world)
 *   every step:  ohmb200_exchange_send(map, own rays ...)    filter, cut, route; returns after queueing
 *                ohmb200_exchange_integrate(map)             waits ON THE DEVICE for every rank's records of the step
 *                                                            (a bounded spin: a missing peer cannot hang the GPU), then
 *                                                            integrates them; returns after queueing
 * Every rank must make the same sequence of steps (a rank with no rays sends an empty batch) with the same ray flags and
 * the same presence of timestamps / intensities.  Occupancy and NDT maps (a TSDF map is sharded with ohmb200_set_partition).
 * With timestamps, set the time base on every rank first (ohmb200_set_first_ray_time).  A single process may drive
 * several maps (same or different devices): call send on all of them, then integrate on all of them.  While the exchange
 * is open the plain integrate calls are refused; ohmb200_exchange_close (or ohmb200_destroy) releases it.
 * world = 1 is valid (the exchange pipeline on one GPU). */
#define OHMB200_EXCHANGE_HANDLE_BYTES 128
typedef struct ohmb200_exchange_handle
{
  unsigned char bytes[OHMB200_EXCHANGE_HANDLE_BYTES];
} ohmb200_exchange_handle;
OHMB200_API int ohmb200_exchange_open(ohmb200_map *map, int rank, int world, size_t max_rays_per_rank,
                                      ohmb200_exchange_handle *handle);
OHMB200_API int ohmb200_exchange_connect(ohmb200_map *map, const ohmb200_exchange_handle *handles, int count);
/* rays etc. in HOST memory (staged like ohmb200_integrate) / already in device memory */
OHMB200_API size_t ohmb200_exchange_send(ohmb200_map *map, const double *rays, size_t element_count,
                                         const float *intensities, const double *timestamps, unsigned ray_flags);
OHMB200_API size_t ohmb200_exchange_send_device(ohmb200_map *map, const double *d_rays, size_t element_count,
                                                const float *d_intensities, const double *d_timestamps,
                                                unsigned ray_flags);
OHMB200_API int ohmb200_exchange_integrate(ohmb200_map *map);
OHMB200_API int ohmb200_exchange_close(ohmb200_map *map);
/* A barrier between the ranks on the device: each rank's stream passes it when every rank's stream has reached it
 * (mailbox flags + a bounded spin; no host wait).  Every rank must call it the same number of times, between steps. */
OHMB200_API int ohmb200_exchange_barrier(ohmb200_map *map);
/* What this rank put into each owner's inbox in the last step (waits for the queued work): segment records (32 bytes
 * each) and sample records (96 bytes, 16 on an occupancy-only map), `world` entries each; either pointer may be NULL.
 * With the per-ray broadcast (64 bytes x own rays x (world - 1) peers; + 60 for NDT maps, + 8 with the traversal
 * layer) this is the step's NVLink traffic — the number a link-bandwidth roofline of the exchange needs. */
OHMB200_API int ohmb200_exchange_last_counts(ohmb200_map *map, uint32_t *segments, uint32_t *samples, int capacity);

/* Per-kernel CUDA-event timing on the map's stream (bench.py's roofline numbers).  When enabled every launch is
 * bracketed by events; ohmb200_kernel_times drains {name -> accumulated ms, launches}. */
#define OHMB200_KERNEL_SLOTS 40
typedef struct ohmb200_kernel_time
{
  char name[32];
  double ms;
  uint64_t launches;
} ohmb200_kernel_time;
OHMB200_API int ohmb200_set_profiling(ohmb200_map *map, int enabled);
OHMB200_API int ohmb200_kernel_times(ohmb200_map *map, ohmb200_kernel_time *out, int capacity, int reset);

/* ---- paging: the GpuLayerCache of this build (ohmgpu/GpuLayerCache.cpp:429-633, GpuCache.h) ------------------------
 * The map is resident in device memory (device_bytes / bytes per region = the slots of the region table).  When a
 * batch may not find enough free slots, the least recently walked regions are copied to a host-side store and their
 * slots freed (GpuLayerCache evicts its oldest cache entry and syncs it to the MapChunk); a region that is touched
 * again is uploaded before the batch updates it (GpuLayerCache::upload on a cache miss).  Reads, enumeration and the
 * region count cover both halves.  Eviction waits for the queued work, so a map that fits never pays for it.
 * ohmb200_set_region_reserve: free slots to guarantee before every batch (default min(4096, capacity / 2)); a single
 * batch that creates more regions than that can still fill the table (ohmb200_sync -> OHMB200_E_CACHE_FULL). */
OHMB200_API int ohmb200_set_region_reserve(ohmb200_map *map, uint32_t free_slots);
/* MapRegionCache::remove (ohm/MapRegionCache.h:52, called by OccupancyMap::cullRegions... OccupancyMap.cpp:678): drops one
 * region — resident or stored — with all its layers.  OHMB200_E_NOT_FOUND if the map does not hold it. */
OHMB200_API int ohmb200_remove_region(ohmb200_map *map, const int16_t key_xyz[3]);
/* resident = regions in device memory, stored = regions in the host store, evicted / paged_in = totals so far. */
OHMB200_API int ohmb200_paging_stats(ohmb200_map *map, uint64_t *resident, uint64_t *stored, uint64_t *evicted,
                                     uint64_t *paged_in);

/* Message of the last failure on this thread. */
OHMB200_API const char *ohmb200_last_error(void);

/* "ohmb200 <version> sm_100a" */
OHMB200_API const char *ohmb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OHMB200_H */
