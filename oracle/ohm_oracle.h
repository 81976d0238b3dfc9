/*
 * ohm_oracle.h — CPU restatement of ohm's single-threaded RayMapper path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the "port" CPU baseline.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (libohmb200.so) never links, loads or calls anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root, csiro-robotics/ohm @ 4e2e769).
 *
 * Pinning status: pinned against the reference's own known-answer tests (tests/test_oracle_kat.py)
 * and against the reference's own compute headers compiled unmodified into oracle/_ref
 * (tests/test_oracle_vs_ref.py, run wherever /root/reference exists).
 */
#ifndef OHM_ORACLE_H
#define OHM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layer ids (also bit positions in oracle_params.layers). Sizes are bytes per voxel.
 * ohm/DefaultLayer.cpp:76-337 */
enum
{
  ORC_LAYER_OCCUPANCY = 0,  /* f32, clear = +inf   (DefaultLayer.cpp:85-91)  */
  ORC_LAYER_MEAN = 1,       /* {u32 coord,u32 count} (VoxelMeanCompute.h:29-33) */
  ORC_LAYER_TRAVERSAL = 2,  /* f32 */
  ORC_LAYER_TOUCH_TIME = 3, /* u32 */
  ORC_LAYER_INCIDENT = 4,   /* u32 */
  ORC_LAYER_COVARIANCE = 5, /* 6 x f32 (CovarianceVoxelCompute.h:56-64) */
  ORC_LAYER_INTENSITY = 6,  /* 2 x f32 */
  ORC_LAYER_HIT_MISS = 7,   /* 2 x u32 */
  ORC_LAYER_TSDF = 8,       /* {f32 weight,f32 distance} (VoxelTsdfCompute.h:20-24) */
  ORC_LAYER_SECONDARY = 9,  /* {f32 m2, u16 range_mean, u16 count} (VoxelSecondarySample.h:29-38) */
  ORC_LAYER_COUNT = 10
};

/* Ray filter kinds: ohm/RayFilter.cpp:15-87 */
enum
{
  ORC_FILTER_NONE = 0,
  ORC_FILTER_GOOD_RAY = 1,  /* goodRayFilter(max_range)  — the OccupancyMap default with 1e10 */
  ORC_FILTER_CLIP_RANGE = 2, /* clipRayFilter(max_length) */
  ORC_FILTER_CLIP_BOX = 3    /* clipBounded(clip_box) (RayFilter.cpp:57-76, Aabb.h:330-450) */
};

typedef struct oracle_params
{
  double resolution;
  int32_t region_dim[3];
  double origin[3];
  float hit_value;
  float miss_value;
  float min_value;
  float max_value;
  float threshold_value;
  int32_t saturate_min;
  int32_t saturate_max;
  uint32_t layers; /* bitset of (1u << ORC_LAYER_*) */
  int32_t filter_kind;
  double filter_range;
  double clip_box[6]; /* ORC_FILTER_CLIP_BOX: min xyz, max xyz */
  /* NDT: ohm/private/NdtMapDetail.h:24-40 */
  float sensor_noise;
  float adaptation_rate;
  float reinit_threshold;
  uint32_t reinit_count;
  uint32_t sample_threshold;
  float initial_intensity_cov;
  int32_t ndt_tm;
  /* TSDF: ohm/VoxelTsdf.h:27-37 */
  float tsdf_max_weight;
  float tsdf_trunc;
  float tsdf_dropoff;
  float tsdf_sparsity;
} oracle_params;

typedef struct oracle_map oracle_map;

/* Statistics gathered while integrating (the units of SURVEY.md §8d's byte model). */
typedef struct oracle_stats
{
  uint64_t rays_in;       /* rays offered */
  uint64_t rays_accepted; /* rays surviving the filter */
  uint64_t voxel_visits;  /* V: walk visits (miss-side updates; TSDF: all visits) */
  uint64_t sample_updates;/* S: explicit sample voxel updates */
} oracle_stats;

void oracle_default_params(oracle_params *p, double resolution);
oracle_map *oracle_map_create(const oracle_params *p);
void oracle_map_destroy(oracle_map *m);
void oracle_map_set_params(oracle_map *m, const oracle_params *p);
void oracle_map_stats(const oracle_map *m, oracle_stats *out);
double oracle_first_ray_time(const oracle_map *m);

/* key = {region x,y,z, local x,y,z}. Returns 0 when the key is null. */
int oracle_voxel_key(const oracle_map *m, const double p[3], int32_t key[6]);
void oracle_voxel_centre(const oracle_map *m, const int32_t key[6], double centre[3]);

/* Walk flags: ohm/LineWalk.h:51-57 */
#define ORC_WALK_EXCLUDE_START 1u
#define ORC_WALK_EXCLUDE_END 2u
/* Returns the number of voxels reported (<= cap written). keys: 6 ints per voxel. */
size_t oracle_walk_segment(const oracle_map *m, const double start[3], const double end[3], unsigned walk_flags,
                           int32_t *keys, double *enter, double *exit, size_t cap);

/* Voxels a walk of every ray reports, computed from the start/end keys only (no walking). */
uint64_t oracle_count_walk_visits(const oracle_map *m, const double *rays, size_t element_count, unsigned walk_flags);

/* RayMapperOccupancy / RayMapperNdt / RayMapperTsdf ::integrateRays. rays = [origin,sample]* as f64 xyz. */
size_t oracle_integrate_occupancy(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                                  const double *timestamps, unsigned ray_flags);
size_t oracle_integrate_ndt(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                            const double *timestamps, unsigned ray_flags);
size_t oracle_integrate_tsdf(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                             const double *timestamps, unsigned ray_flags);

/* ohm::RaysQuery::onExecute (ohm/RaysQuery.cpp:109-199): per ray, walk until the first occupied voxel.  ranges[n] = exit
 * range of the last voxel that is not occupied (a float, stored as double), unobserved_volumes[n] = volume_coefficient * sum over
 * unobserved voxels of (exit^3 - enter^3), terminal_states[n] = OccupancyType of the last voxel visited (ohm/OccupancyType.h:
 * -2 null, -1 unobserved, 0 free, 1 occupied), terminal_keys[6n..] = its key.  A ray the filter rejects reports 0, 0,
 * null.  (terminal state/key are NOT reset between rays in the reference: a ray that visits nothing repeats the
 * previous ray's.)  Returns the number of rays. */
#define ORC_OCCUPANCY_NULL (-2)
#define ORC_OCCUPANCY_UNOBSERVED (-1)
#define ORC_OCCUPANCY_FREE 0
#define ORC_OCCUPANCY_OCCUPIED 1
size_t oracle_rays_query(const oracle_map *m, const double *rays, size_t element_count, double volume_coefficient,
                         double *ranges, double *unobserved_volumes, int *terminal_states, int32_t *terminal_keys);

/* RayMapperSecondarySample::integrateRays (ohm/RayMapperSecondarySample.cpp:37-74): for every ray, Welford update of
 * the secondary-sample voxel at the END point with range = |end - start| (VoxelSecondarySample.h:87-99).  No ray
 * filter, no walk.  The map must hold ORC_LAYER_SECONDARY. */
size_t oracle_integrate_secondary(oracle_map *m, const double *rays, size_t element_count);

size_t oracle_region_count(const oracle_map *m);
/* Writes up to cap region keys (3 x int16 each), sorted (z,y,x ascending); returns the total count. */
size_t oracle_region_keys(const oracle_map *m, int16_t *keys, size_t cap);
/* Pointer to the layer's voxel array for a region (x + y*dx + z*dx*dy order), or NULL. */
const void *oracle_region_layer(const oracle_map *m, const int16_t key[3], int layer);
size_t oracle_layer_voxel_bytes(int layer);

/* Scalar helpers exposed for known-answer tests. */
uint32_t oracle_sub_voxel_update(uint32_t coord, uint32_t count, const double local[3], double resolution);
void oracle_sub_voxel_to_local(uint32_t coord, double resolution, double out[3]);
uint32_t oracle_update_incident_normal(uint32_t packed, const float incident[3], uint32_t count);
void oracle_decode_normal(uint32_t packed, float out[3]);
uint32_t oracle_encode_normal(const float n[3]);
uint32_t oracle_encode_touch_time(double timebase, double timestamp);
int oracle_calculate_hit_with_covariance(float cov[6], float *value, const double sample[3], const double mean[3],
                                         uint32_t count, float hit_value, float uninit, float resolution,
                                         float reinit_threshold, uint32_t reinit_count);
void oracle_calculate_miss_ndt(const float cov[6], float *value, int *is_miss, const double sensor[3],
                               const double sample[3], const double mean[3], uint32_t count, float uninit,
                               float miss_value, float adaptation_rate, float sensor_noise, uint32_t sample_threshold);
int oracle_calculate_tsdf(const double sensor[3], const double sample[3], const double centre[3], float trunc,
                          float max_weight, float dropoff, float sparsity, float *weight, float *distance);
void oracle_occupancy_adjust_miss(float *v, float initial, float adj, float uninit, float min_value, float sat_min,
                                  float sat_max, int null_update);
void oracle_occupancy_adjust_hit(float *v, float initial, float adj, float uninit, float max_value, float sat_min,
                                 float sat_max, int null_update);

#ifdef __cplusplus
}
#endif
#endif
