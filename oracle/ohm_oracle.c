/*
 * ohm_oracle.c — plain-C restatement of ohm's CPU ray integration (the parity oracle).
 *
 * TEST INFRASTRUCTURE ONLY (see ohm_oracle.h).  Build: gcc -O2 -ffp-contract=off -fno-fast-math.
 * The reference CPU build (x86-64, no -march) never contracts to FMA, so neither may we.
 *
 * Restated from (reference file:line):
 *   key maths            ohm/MapCoord.h:32-93, ohm/MapRegion.cpp:32-69, ohm/OccupancyMap.h:757-778,827-846,887-901
 *   line walk            ohm/LineWalkCompute.h:162-413, ohm/LineWalk.h:59-129
 *   occupancy update     ohm/VoxelOccupancyCompute.h:44-153, ohm/RayMapperOccupancy.cpp:68-339
 *   voxel mean           ohm/VoxelMeanCompute.h:69-152
 *   incident normal      ohm/VoxelIncidentCompute.h:35-112
 *   touch time           ohm/VoxelTouchTimeCompute.h:18-27
 *   NDT                  ohm/CovarianceVoxelCompute.h:90-635, ohm/RayMapperNdt.cpp:84-407
 *   TSDF                 ohm/VoxelTsdfCompute.h:57-136, ohm/RayMapperTsdf.cpp:87-182
 *   ray filters          ohm/RayFilter.cpp:15-55
 * GLM (un-vendored, unpinned dependency) contributes only dot/length/normalize on dvec3; restated as
 *   dot(a,b) = (ax*bx + ay*by) + az*bz,  normalize(v) = v * (1.0 / sqrt(dot(v,v)))   [GLM 0.9.9 generic path].
 */
#include "ohm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Map container                                                                               */
/* ------------------------------------------------------------------------------------------ */
static const size_t k_layer_bytes[ORC_LAYER_COUNT] = { 4, 8, 4, 4, 4, 24, 8, 8, 8, 8 };

typedef struct chunk
{
  int16_t region[3];
  void *layers[ORC_LAYER_COUNT];
} chunk;

struct oracle_map
{
  oracle_params p;
  double region_size[3]; /* region_spatial_dimensions (OccupancyMap.cpp:204-206) */
  size_t voxels_per_region;
  chunk **table; /* open addressing, NULL = empty */
  size_t table_cap;
  size_t region_count;
  double first_ray_time; /* < 0 = unset (OccupancyMap.cpp:343-347) */
  oracle_stats stats;
  chunk *last_chunk;
};

size_t oracle_layer_voxel_bytes(int layer)
{
  return (layer >= 0 && layer < ORC_LAYER_COUNT) ? k_layer_bytes[layer] : 0;
}

void oracle_default_params(oracle_params *p, double resolution)
{
  /* ohm/OccupancyMap.cpp:195-213, ohm/OccupancyMap.h:24-26, NdtMapDetail.h:24-40, NdtMap.cpp:33-45, VoxelTsdf.h:27-37 */
  memset(p, 0, sizeof(*p));
  p->resolution = resolution;
  p->region_dim[0] = p->region_dim[1] = p->region_dim[2] = 32;
  p->hit_value = logf(0.9f / (1.0f - 0.9f));
  p->miss_value = logf(0.45f / (1.0f - 0.45f));
  p->min_value = -2.0f;
  p->max_value = 3.511f;
  p->threshold_value = logf(0.5f / (1.0f - 0.5f));
  p->layers = 1u << ORC_LAYER_OCCUPANCY;
  p->filter_kind = ORC_FILTER_GOOD_RAY;
  p->filter_range = 1e10;
  p->sensor_noise = 0.05f;
  p->adaptation_rate = 0.2f; /* NdtMap::ndtAdaptationRateFromMissProbability(0.45) (NdtMap.cpp:194-213) */
  p->reinit_threshold = logf(0.2f / (1.0f - 0.2f));
  p->reinit_count = 100;
  p->sample_threshold = 3;
  p->initial_intensity_cov = 1.0f;
  p->tsdf_max_weight = 1e4f;
  p->tsdf_trunc = 0.1f;
  p->tsdf_dropoff = 0.0f;
  p->tsdf_sparsity = 1.0f;
}

oracle_map *oracle_map_create(const oracle_params *p)
{
  oracle_map *m = (oracle_map *)calloc(1, sizeof(oracle_map));
  m->p = *p;
  for (int a = 0; a < 3; ++a)
  {
    m->region_size[a] = p->region_dim[a] * p->resolution;
  }
  m->voxels_per_region = (size_t)p->region_dim[0] * p->region_dim[1] * p->region_dim[2];
  m->table_cap = 1024;
  m->table = (chunk **)calloc(m->table_cap, sizeof(chunk *));
  m->first_ray_time = -1.0;
  return m;
}

static void chunk_free(chunk *c)
{
  for (int l = 0; l < ORC_LAYER_COUNT; ++l)
  {
    free(c->layers[l]);
  }
  free(c);
}

void oracle_map_destroy(oracle_map *m)
{
  if (!m)
  {
    return;
  }
  for (size_t i = 0; i < m->table_cap; ++i)
  {
    if (m->table[i])
    {
      chunk_free(m->table[i]);
    }
  }
  free(m->table);
  free(m);
}

void oracle_map_set_params(oracle_map *m, const oracle_params *p)
{
  /* Mutable parameters only (OccupancyMap::setHitValue/setMissValue/..., NdtMap setters); geometry is fixed. */
  const uint32_t layers = m->p.layers;
  const double res = m->p.resolution;
  int32_t dim[3];
  memcpy(dim, m->p.region_dim, sizeof(dim));
  m->p = *p;
  m->p.layers = layers;
  m->p.resolution = res;
  memcpy(m->p.region_dim, dim, sizeof(dim));
}

void oracle_map_stats(const oracle_map *m, oracle_stats *out)
{
  *out = m->stats;
}

double oracle_first_ray_time(const oracle_map *m)
{
  return m->first_ray_time;
}

static uint64_t region_hash(const int16_t r[3])
{
  uint64_t k = ((uint64_t)(uint16_t)r[0]) | ((uint64_t)(uint16_t)r[1] << 16) | ((uint64_t)(uint16_t)r[2] << 32);
  k *= 0x9E3779B97F4A7C15ull;
  return k ^ (k >> 29);
}

static chunk *chunk_create(const oracle_map *m, const int16_t r[3])
{
  chunk *c = (chunk *)calloc(1, sizeof(chunk));
  memcpy(c->region, r, sizeof(c->region));
  for (int l = 0; l < ORC_LAYER_COUNT; ++l)
  {
    if (m->p.layers & (1u << l))
    {
      c->layers[l] = calloc(m->voxels_per_region, k_layer_bytes[l]);
    }
  }
  if (c->layers[ORC_LAYER_OCCUPANCY])
  {
    /* Unobserved = +inf (DefaultLayer.cpp:87-91). */
    float *occ = (float *)c->layers[ORC_LAYER_OCCUPANCY];
    for (size_t i = 0; i < m->voxels_per_region; ++i)
    {
      occ[i] = INFINITY;
    }
  }
  return c;
}

static void table_insert(chunk **table, size_t cap, chunk *c)
{
  size_t i = region_hash(c->region) & (cap - 1);
  while (table[i])
  {
    i = (i + 1) & (cap - 1);
  }
  table[i] = c;
}

static chunk *map_region(oracle_map *m, const int16_t r[3], int create)
{
  if (m->last_chunk && m->last_chunk->region[0] == r[0] && m->last_chunk->region[1] == r[1] &&
      m->last_chunk->region[2] == r[2])
  {
    return m->last_chunk;
  }
  size_t i = region_hash(r) & (m->table_cap - 1);
  while (m->table[i])
  {
    chunk *c = m->table[i];
    if (c->region[0] == r[0] && c->region[1] == r[1] && c->region[2] == r[2])
    {
      m->last_chunk = c;
      return c;
    }
    i = (i + 1) & (m->table_cap - 1);
  }
  if (!create)
  {
    return NULL;
  }
  if ((m->region_count + 1) * 2 > m->table_cap)
  {
    size_t ncap = m->table_cap * 2;
    chunk **nt = (chunk **)calloc(ncap, sizeof(chunk *));
    for (size_t j = 0; j < m->table_cap; ++j)
    {
      if (m->table[j])
      {
        table_insert(nt, ncap, m->table[j]);
      }
    }
    free(m->table);
    m->table = nt;
    m->table_cap = ncap;
  }
  chunk *c = chunk_create(m, r);
  table_insert(m->table, m->table_cap, c);
  ++m->region_count;
  m->last_chunk = c;
  return c;
}

size_t oracle_region_count(const oracle_map *m)
{
  return m->region_count;
}

static int cmp_region(const void *a, const void *b)
{
  const int16_t *ra = (const int16_t *)a;
  const int16_t *rb = (const int16_t *)b;
  for (int i = 2; i >= 0; --i)
  {
    if (ra[i] != rb[i])
    {
      return (ra[i] < rb[i]) ? -1 : 1;
    }
  }
  return 0;
}

size_t oracle_region_keys(const oracle_map *m, int16_t *keys, size_t cap)
{
  int16_t *all = (int16_t *)malloc(sizeof(int16_t) * 3 * (m->region_count + 1));
  size_t n = 0;
  for (size_t i = 0; i < m->table_cap; ++i)
  {
    if (m->table[i])
    {
      memcpy(all + 3 * n, m->table[i]->region, sizeof(int16_t) * 3);
      ++n;
    }
  }
  qsort(all, n, sizeof(int16_t) * 3, cmp_region);
  memcpy(keys, all, sizeof(int16_t) * 3 * (n < cap ? n : cap));
  free(all);
  return n;
}

const void *oracle_region_layer(const oracle_map *m, const int16_t key[3], int layer)
{
  if (layer < 0 || layer >= ORC_LAYER_COUNT)
  {
    return NULL;
  }
  chunk *c = map_region((oracle_map *)m, key, 0);
  return c ? c->layers[layer] : NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* Key maths                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* ohm/MapCoord.h:85-93 pointToRegionCoord<double> */
static int point_to_region_coord(double coord, double resolution)
{
  return (int)floor(coord / resolution + (double)0.5f);
}

/* ohm/MapCoord.h:41-80 pointToRegionVoxel<double> */
static int point_to_region_voxel(double coord, double voxel_resolution, double region_resolution)
{
  const double epsilon = (double)1e-6f;
  if (-epsilon <= coord && coord < 0)
  {
    coord = 0;
  }
  else if (coord >= region_resolution && coord - epsilon < region_resolution)
  {
    coord -= epsilon;
  }
  return (int)floor(coord / voxel_resolution);
}

/* ohm/OccupancyMap.cpp:859-886 voxelKey -> ohm/MapRegion.cpp:32-69 */
int oracle_voxel_key(const oracle_map *m, const double p[3], int32_t key[6])
{
  int ok = 1;
  for (int a = 0; a < 3; ++a)
  {
    const int16_t rc = (int16_t)point_to_region_coord(p[a] - m->p.origin[a], m->region_size[a]);
    const double centre = rc * m->region_size[a]; /* regionCentreCoord, MapCoord.h:36-39 */
    const double region_min = centre - 0.5 * m->region_size[a];
    const double local = p[a] - m->p.origin[a] - region_min;
    const int q = point_to_region_voxel(local, m->p.resolution, m->region_size[a]);
    key[a] = rc;
    key[3 + a] = q;
    ok = ok && (0 <= q && q < m->p.region_dim[a]);
  }
  if (!ok)
  {
    /* Key::kNull */
    key[0] = key[1] = key[2] = INT16_MIN;
    key[3] = key[4] = key[5] = 255;
  }
  return ok;
}

/* ohm/OccupancyMap.h:757-778 voxelCentre (region key passes through float, exact for int16) */
void oracle_voxel_centre(const oracle_map *m, const int32_t key[6], double centre[3])
{
  for (int a = 0; a < 3; ++a)
  {
    double c = (double)(float)key[a];
    c *= m->region_size[a];
    c -= 0.5 * m->region_size[a];
    c += m->p.origin[a];
    c += (double)key[3 + a] * m->p.resolution;
    c += 0.5 * m->p.resolution;
    centre[a] = c;
  }
}

/* ohm/OccupancyMap.h:827-846 stepKey */
static void step_key(const oracle_map *m, int32_t key[6], int axis, int dir)
{
  int local = key[3 + axis] + dir;
  int region = key[axis];
  if (local < 0)
  {
    --region;
    local = m->p.region_dim[axis] - 1;
  }
  else if (local >= m->p.region_dim[axis])
  {
    ++region;
    local = 0;
  }
  key[3 + axis] = (uint8_t)local;
  key[axis] = (int16_t)(uint16_t)region;
}

static int keys_equal(const int32_t a[6], const int32_t b[6])
{
  return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3] && a[4] == b[4] && a[5] == b[5];
}

/* ------------------------------------------------------------------------------------------ */
/* Line walk: ohm/LineWalkCompute.h                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef int (*visit_fn)(void *ctx, const int32_t key[6], double enter_range, double exit_range);

typedef struct walk_steps
{
  double time_next[3];
  double initial_delta[3];
  double step_delta[3];
  int sign[3];
  double length;
} walk_steps;

static int step_dir(int sign)
{
  return -2 * sign + 1; /* LineWalkCompute.h:162-173 */
}

/* LineWalkCompute.h:177-186 */
static void wall_exit(const double origin[3], const double inv[3], const int sign[3], const double vmin[3],
                      const double vmax[3], double out[3])
{
  for (int a = 0; a < 3; ++a)
  {
    const double bound = (1 - sign[a]) ? vmax[a] : vmin[a];
    out[a] = (bound - origin[a]) * inv[a];
  }
}

/* LineWalkCompute.h:188-280 walkInitRay + walkCalculateSteps */
static void calculate_steps(walk_steps *ws, const double start[3], const double end[3], const double centre[3],
                            double res, double length_epsilon)
{
  double dir[3], inv[3], vmin[3], vmax[3], delta[3];
  for (int a = 0; a < 3; ++a)
  {
    dir[a] = end[a] - start[a];
  }
  double length = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
  length = (length > length_epsilon) ? sqrt(length) : 0;
  for (int a = 0; a < 3; ++a)
  {
    ws->sign[a] = dir[a] < 0;
  }
  for (int a = 0; a < 3; ++a)
  {
    dir[a] /= length;
  }
  for (int a = 0; a < 3; ++a)
  {
    inv[a] = (length > 0) ? 1 / dir[a] : 0;
  }
  for (int a = 0; a < 3; ++a)
  {
    vmin[a] = centre[a] - 0.5 * res;
    vmax[a] = centre[a] + 0.5 * res;
  }
  wall_exit(start, inv, ws->sign, vmin, vmax, ws->initial_delta);
  for (int a = 0; a < 3; ++a)
  {
    const double shift = step_dir(ws->sign[a]) * res;
    vmin[a] += shift;
    vmax[a] += shift;
  }
  wall_exit(start, inv, ws->sign, vmin, vmax, delta);
  for (int a = 0; a < 3; ++a)
  {
    if (delta[a] != INFINITY)
    {
      delta[a] -= ws->initial_delta[a];
    }
    ws->step_delta[a] = delta[a];
    ws->time_next[a] = ws->initial_delta[a];
  }
  ws->length = length;
}

/* LineWalkCompute.h:282-289 */
static int select_next_axis(const double t[3])
{
  int axis = 0;
  axis = (t[axis] < t[1]) ? axis : 1;
  axis = (t[axis] < t[2]) ? axis : 2;
  return axis;
}

/* LineWalkCompute.h:291-307 */
static unsigned step_next(const oracle_map *m, walk_steps *ws, int32_t key[6], int *axis, int remaining[3],
                          int stepped[3])
{
  const int a = *axis;
  const int dir = step_dir(ws->sign[a]);
  step_key(m, key, a, dir);
  remaining[a] -= dir;
  stepped[a] += dir;
  ws->time_next[a] = remaining[a] ? ws->initial_delta[a] + ws->step_delta[a] * abs(stepped[a]) : INFINITY;
  const unsigned change = (remaining[a] == 0) ? (1u << a) : 0u;
  *axis = select_next_axis(ws->time_next);
  return change;
}

/* LineWalkCompute.h:345-413 walkLineVoxels */
static unsigned walk_line(const oracle_map *m, const double start[3], const double end[3], const int32_t skey[6],
                          const int32_t ekey[6], const double start_centre[3], unsigned flags, double length_epsilon,
                          visit_fn visit, void *ctx)
{
  walk_steps ws;
  calculate_steps(&ws, start, end, start_centre, m->p.resolution, length_epsilon);

  int remaining[3];
  int stepped[3] = { 0, 0, 0 };
  /* walkKeyDiff = rangeBetween(start, end) (LineWalk.h:63-70, OccupancyMap.h:887-901) */
  for (int a = 0; a < 3; ++a)
  {
    remaining[a] = ekey[3 + a] - skey[3 + a] + (int16_t)(ekey[a] - skey[a]) * m->p.region_dim[a];
  }

  int32_t cur[6];
  memcpy(cur, skey, sizeof(cur));
  double last_time = 0;
  int axis = 0;
  unsigned count = 0;
  unsigned limit = 0;
  int cont = 1;
  for (int a = 0; a < 3; ++a)
  {
    limit |= (remaining[a] == 0) ? (1u << a) : 0u;
    ws.time_next[a] = remaining[a] ? ws.initial_delta[a] : INFINITY;
  }
  axis = select_next_axis(ws.time_next);

  if (flags & ORC_WALK_EXCLUDE_START)
  {
    last_time = ws.time_next[axis];
    ++count;
    limit |= step_next(m, &ws, cur, &axis, remaining, stepped);
  }

  while (cont && limit < 7u && !keys_equal(cur, ekey))
  {
    cont = visit(ctx, cur, last_time, ws.time_next[axis]);
    last_time = ws.time_next[axis];
    ++count;
    limit |= step_next(m, &ws, cur, &axis, remaining, stepped);
  }

  if (cont && (flags & ORC_WALK_EXCLUDE_END) == 0u)
  {
    visit(ctx, ekey, last_time, ws.length);
    ++count;
  }
  return count;
}

/* ohm/LineWalk.h:112-129 walkSegmentKeys */
static unsigned walk_segment_keys(const oracle_map *m, const double start[3], const double end[3], unsigned flags,
                                  visit_fn visit, void *ctx)
{
  int32_t skey[6], ekey[6];
  if (!oracle_voxel_key(m, start, skey) || !oracle_voxel_key(m, end, ekey))
  {
    return 0;
  }
  double centre[3];
  oracle_voxel_centre(m, skey, centre);
  return walk_line(m, start, end, skey, ekey, centre, flags, 1e-6, visit, ctx);
}

typedef struct collect_ctx
{
  int32_t *keys;
  double *enter;
  double *exit;
  size_t cap;
  size_t n;
} collect_ctx;

static int collect_visit(void *vctx, const int32_t key[6], double enter_range, double exit_range)
{
  collect_ctx *c = (collect_ctx *)vctx;
  if (c->n < c->cap)
  {
    if (c->keys)
    {
      memcpy(c->keys + 6 * c->n, key, sizeof(int32_t) * 6);
    }
    if (c->enter)
    {
      c->enter[c->n] = enter_range;
    }
    if (c->exit)
    {
      c->exit[c->n] = exit_range;
    }
  }
  ++c->n;
  return 1;
}

size_t oracle_walk_segment(const oracle_map *m, const double start[3], const double end[3], unsigned walk_flags,
                           int32_t *keys, double *enter, double *exit, size_t cap)
{
  collect_ctx c = { keys, enter, exit, cap, 0 };
  walk_segment_keys(m, start, end, walk_flags, collect_visit, &c);
  return c.n;
}

/* Number of voxels walkSegmentKeys reports for each ray, from the keys alone: 1 + |dx|+|dy|+|dz| minus the excluded
 * start/end voxels (LineWalkCompute.h:354-410).  An independent check of V for runs too large to walk on the CPU. */
uint64_t oracle_count_walk_visits(const oracle_map *m, const double *rays, size_t element_count, unsigned walk_flags)
{
  uint64_t total = 0;
  for (size_t i = 0; i + 1 < element_count; i += 2)
  {
    int32_t s[6], e[6];
    if (!oracle_voxel_key(m, rays + 3 * i, s) || !oracle_voxel_key(m, rays + 3 * i + 3, e))
    {
      continue;
    }
    int64_t steps = 0;
    for (int a = 0; a < 3; ++a)
    {
      const int64_t d = (int64_t)(e[3 + a] - s[3 + a]) + (int64_t)(int16_t)(e[a] - s[a]) * m->p.region_dim[a];
      steps += d < 0 ? -d : d;
    }
    if (steps == 0)
    {
      total += (walk_flags & ORC_WALK_EXCLUDE_END) ? 0u : 1u;
      continue;
    }
    int64_t visits = steps + 1;
    visits -= (walk_flags & ORC_WALK_EXCLUDE_START) ? 1 : 0;
    visits -= (walk_flags & ORC_WALK_EXCLUDE_END) ? 1 : 0;
    total += (uint64_t)visits;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------ */
/* Ray filters: ohm/RayFilter.cpp:15-55                                                        */
/* ------------------------------------------------------------------------------------------ */
#define RFF_INVALID 1u
#define RFF_CLIPPED_START 2u
#define RFF_CLIPPED_END 4u

static int vec_finite(const double v[3])
{
  return !(isnan(v[0]) || isnan(v[1]) || isnan(v[2])) && !(isinf(v[0]) || isinf(v[1]) || isinf(v[2]));
}

static double dot3(const double a[3], const double b[3])
{
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

static int apply_filter(const oracle_map *m, double start[3], double end[3], unsigned *filter_flags)
{
  const double range = m->p.filter_range;
  switch (m->p.filter_kind)
  {
  default:
  case ORC_FILTER_NONE:
    return 1;
  case ORC_FILTER_GOOD_RAY: {
    int good = vec_finite(start) && vec_finite(end);
    const double ray[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
    good = good && (range <= 0 || dot3(ray, ray) <= range * range);
    if (!good)
    {
      *filter_flags |= RFF_INVALID;
    }
    return good;
  }
  case ORC_FILTER_CLIP_BOX: {
    /* clipBounded (RayFilter.cpp:57-76) -> Aabb::clipLine / rayIntersect / contains (Aabb.h:309-450) */
    const double *lo = m->p.clip_box, *hi = m->p.clip_box + 3;
    const double origin[3] = { start[0], start[1], start[2] };
    double dir[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
    unsigned clip = 0;
    int clipped = 0;
    if (!(dot3(dir, dir) < 1e-9))
    {
      const double length = sqrt(dot3(dir, dir));
      double inv[3], t0, t1, tmin, tmax;
      int sign[3], miss;
      for (int a = 0; a < 3; ++a)
      {
        dir[a] /= length;
        inv[a] = 1.0 / dir[a];
        sign[a] = dir[a] < 0.0;
      }
      t0 = ((sign[0] ? hi[0] : lo[0]) - origin[0]) * inv[0];
      t1 = ((sign[0] ? lo[0] : hi[0]) - origin[0]) * inv[0];
      tmin = ((sign[1] ? hi[1] : lo[1]) - origin[1]) * inv[1];
      tmax = ((sign[1] ? lo[1] : hi[1]) - origin[1]) * inv[1];
      miss = (t0 > tmax) + (tmin > t1);
      t0 = (tmin > t0 || isnan(t0)) ? tmin : t0;
      t1 = (tmax < t1 || isnan(t1)) ? tmax : t1;
      tmin = ((sign[2] ? hi[2] : lo[2]) - origin[2]) * inv[2];
      tmax = ((sign[2] ? lo[2] : hi[2]) - origin[2]) * inv[2];
      miss = !!((t0 > tmax) + (tmin > t1) + !!miss);
      t0 = (tmin > t0 || isnan(t0)) ? tmin : t0;
      t1 = (tmax < t1 || isnan(t1)) ? tmax : t1;
      if (!miss)
      {
        if (t0 > 0 && t0 < length)
        {
          for (int a = 0; a < 3; ++a)
          {
            start[a] = origin[a] + dir[a] * t0;
          }
          clip |= 1u;
          ++clipped;
        }
        if (t1 > 0 && t1 < length)
        {
          for (int a = 0; a < 3; ++a)
          {
            end[a] = origin[a] + dir[a] * t1;
          }
          clip |= 2u;
          ++clipped;
        }
      }
    }
    if (clipped)
    {
      int s_in = 1, e_in = 1;
      for (int a = 0; a < 3; ++a)
      {
        s_in = s_in && !(hi[a] < start[a]) && !(lo[a] > start[a]);
        e_in = e_in && !(hi[a] < end[a]) && !(lo[a] > end[a]);
      }
      if (!s_in && !e_in)
      {
        return 0;
      }
    }
    *filter_flags |= (clip & 1u) ? RFF_CLIPPED_START : 0u;
    *filter_flags |= (clip & 2u) ? RFF_CLIPPED_END : 0u;
    return 1;
  }
  case ORC_FILTER_CLIP_RANGE: {
    const int good = vec_finite(start) && vec_finite(end);
    double ray[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
    const double len2 = dot3(ray, ray);
    if (good && range > 0 && len2 > range * range)
    {
      const double len = sqrt(len2);
      for (int a = 0; a < 3; ++a)
      {
        ray[a] /= len;
        end[a] = start[a] + ray[a] * range;
      }
      *filter_flags |= RFF_CLIPPED_END;
    }
    if (!good)
    {
      *filter_flags |= RFF_INVALID;
    }
    return good;
  }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Per-voxel arithmetic                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* ohm/VoxelOccupancyCompute.h:110-120 */
void oracle_occupancy_adjust_miss(float *v, float initial, float adj, float uninit, float min_value, float sat_min,
                                  float sat_max, int null_update)
{
  const int uninitialised = initial == uninit;
  const float base = (null_update || !uninitialised) ? initial : 0.0f;
  adj = (!null_update && (uninitialised || (sat_min < initial && initial < sat_max))) ? adj : 0.0f;
  *v = (base != uninit) ? fmaxf(min_value, base + adj) : base;
}

/* ohm/VoxelOccupancyCompute.h:44-54 */
void oracle_occupancy_adjust_hit(float *v, float initial, float adj, float uninit, float max_value, float sat_min,
                                 float sat_max, int null_update)
{
  const int uninitialised = initial == uninit;
  const float base = (null_update || !uninitialised) ? initial : 0.0f;
  adj = (!null_update && (uninitialised || (sat_min < initial && initial < sat_max))) ? adj : 0.0f;
  *v = (base != uninit) ? fminf(base + adj, max_value) : base;
}

/* ohm/VoxelOccupancyCompute.h:75-85 */
static void occupancy_adjust_up(float *v, float initial, float adjusted, float uninit, float max_value, float sat_min,
                                float sat_max, int null_update)
{
  const int uninitialised = initial == uninit;
  adjusted = (!null_update && (uninitialised || (sat_min < initial && initial < sat_max))) ? adjusted : initial;
  *v = (adjusted != uninit) ? fminf(max_value, adjusted) : adjusted;
}

/* ohm/VoxelOccupancyCompute.h:144-153 */
static void occupancy_adjust_down(float *v, float initial, float adjusted, float uninit, float min_value,
                                  float sat_min, float sat_max, int null_update)
{
  const int uninitialised = initial == uninit;
  adjusted = (!null_update && (uninitialised || (sat_min < initial && initial < sat_max))) ? adjusted : initial;
  *v = (adjusted != uninit) ? fmaxf(min_value, adjusted) : adjusted;
}

/* ohm/VoxelMeanCompute.h:69-92 subVoxelCoord<dvec3,double> */
static uint32_t sub_voxel_coord(const double local[3], double resolution)
{
  const int mean_positions = (1 << 10) - 1;
  const double mean_resolution = resolution / (double)mean_positions;
  const double offset = (double)0.5f * resolution;
  uint32_t pattern = 0;
  for (int a = 0; a < 3; ++a)
  {
    int pos = point_to_region_coord(local[a] + offset, mean_resolution);
    pos = (pos >= 0 ? (pos < (1 << 10) ? pos : mean_positions) : 0);
    pattern |= ((uint32_t)pos) << (10 * a);
  }
  return pattern | (1u << 31);
}

/* ohm/VoxelMeanCompute.h:102-122 subVoxelToLocalCoord<dvec3> (the used-bit test there is constant-true) */
void oracle_sub_voxel_to_local(uint32_t coord, double resolution, double out[3])
{
  const int mean_positions = (1 << 10) - 1;
  const double mean_resolution = resolution / (double)mean_positions;
  const double offset = (double)0.5f * resolution;
  for (int a = 0; a < 3; ++a)
  {
    out[a] = (int)((coord >> (10 * a)) & (uint32_t)mean_positions) * mean_resolution - offset;
  }
}

/* ohm/VoxelMeanCompute.h:134-152 subVoxelUpdate<dvec3,double> */
uint32_t oracle_sub_voxel_update(uint32_t coord, uint32_t count, const double local[3], double resolution)
{
  double mean[3];
  oracle_sub_voxel_to_local(coord, resolution, mean);
  const double one_on_count_plus_one = (double)1 / (double)(count + 1u);
  for (int a = 0; a < 3; ++a)
  {
    mean[a] += (local[a] - mean[a]) * one_on_count_plus_one;
  }
  return sub_voxel_coord(mean, resolution);
}

/* ohm/VoxelIncidentCompute.h:35-54 */
void oracle_decode_normal(uint32_t packed, float n[3])
{
  n[0] = (2.0f * ((float)((packed >> 0) & 0x3FFFu) / 16383.0f)) - 1.0f;
  n[1] = (2.0f * ((float)((packed >> 15) & 0x3FFFu) / 16383.0f)) - 1.0f;
  n[0] = fmaxf(-1.0f, fminf(n[0], 1.0f));
  n[1] = fmaxf(-1.0f, fminf(n[1], 1.0f));
  n[2] = fmaxf(-1.0f, fminf(1.0f - (n[0] * n[0] + n[1] * n[1]), 1.0f));
  const int set = (packed & (1u << 30)) != 0;
  n[0] = set ? n[0] : 0.0f;
  n[1] = set ? n[1] : 0.0f;
  n[2] = set ? sqrtf(n[2]) : 0.0f;
  n[2] *= (packed & (1u << 31)) ? -1.0f : 1.0f;
}

/* ohm/VoxelIncidentCompute.h:69-91 */
uint32_t oracle_encode_normal(const float normal_in[3])
{
  uint32_t n = 0;
  const float x = 0.5f * (fmaxf(-1.0f, fminf(normal_in[0], 1.0f)) + 1.0f);
  const float y = 0.5f * (fmaxf(-1.0f, fminf(normal_in[1], 1.0f)) + 1.0f);
  uint32_t i = (uint32_t)(x * 16383.0f);
  n |= (i & 0x3FFFu) << 0;
  i = (uint32_t)(y * 16383.0f);
  n |= (i & 0x3FFFu) << 15;
  n &= ~((1u << 30) | (1u << 31));
  n |= (normal_in[2] < 0) ? (1u << 31) : 0u;
  /* the set test reads the *remapped* x,y (VoxelIncidentCompute.h:88) */
  n |= (x != 0.0f || y != 0.0f || normal_in[2] != 0.0f) ? (1u << 30) : 0u;
  return n;
}

/* ohm/VoxelIncidentCompute.h:93-112 */
uint32_t oracle_update_incident_normal(uint32_t packed, const float incident_in[3], uint32_t count)
{
  float n[3];
  float inc[3] = { incident_in[0], incident_in[1], incident_in[2] };
  oracle_decode_normal(packed, n);
  count = ((n[0] != 0 || n[1] != 0 || n[2] != 0) && count) ? count : 0;
  const float one_on_count_plus_one = 1.0f / (float)(count + 1u);
  float len2 = inc[0] * inc[0] + inc[1] * inc[1] + inc[2] * inc[2];
  /* `1.0f / sqrt(float)` in the reference resolves to the double sqrt (only <cmath>'s ::sqrt(double) is visible in
   * namespace ohm), so the reciprocal is formed in double and narrowed by glm's `vec *= scalar`; verified against
   * the reference itself (oracle/_ref, tests/test_oracle_vs_ref.py). */
  float s = (len2 > 1e-6f) ? (float)(1.0 / sqrt((double)len2)) : 0.0f;
  for (int a = 0; a < 3; ++a)
  {
    inc[a] *= s;
  }
  for (int a = 0; a < 3; ++a)
  {
    n[a] += (inc[a] - n[a]) * one_on_count_plus_one;
  }
  len2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  s = (len2 > 1e-6f) ? (float)(1.0 / sqrt((double)len2)) : 0.0f;
  for (int a = 0; a < 3; ++a)
  {
    n[a] *= s;
  }
  return oracle_encode_normal(n);
}

/* ohm/VoxelTouchTimeCompute.h:18-27 */
uint32_t oracle_encode_touch_time(double timebase, double timestamp)
{
  /* gcc/x86-64 lowers (unsigned)double to a 64-bit truncating convert; made explicit here */
  return (uint32_t)(int64_t)((timestamp - timebase) / 0.001);
}

/* ------------------------------------------------------------------------------------------ */
/* NDT: ohm/CovarianceVoxelCompute.h                                                           */
/* ------------------------------------------------------------------------------------------ */

/* :101-117 packedDot */
static double packed_dot(const double A[9], int j, int k)
{
  static const int col_first_el[3] = { 0, 1, 3 };
  const int indj = col_first_el[j];
  const int indk = col_first_el[k];
  const int mm = (j <= k) ? j : k;
  double d = A[6 + k] * A[6 + j];
  for (int i = 0; i <= mm; ++i)
  {
    d += A[indj + i] * A[indk + i];
  }
  return d;
}

/* :183-204 solveTriangular */
static void solve_triangular(const float cov[6], const double y[3], double x[3])
{
  double d;
  d = y[0];
  x[0] = d / cov[0];
  d = y[1];
  d -= cov[1] * x[0];
  x[1] = d / cov[2];
  d = y[2];
  d -= cov[3] * x[0];
  d -= cov[4] * x[1];
  x[2] = d / cov[5];
}

/* :227-267 calculateSampleLikelihoods */
static void sample_likelihoods(const float cov[6], const double sensor[3], const double sample[3],
                               const double mean[3], float sensor_noise, double *p_voxel, double *p_sample,
                               double x_ml[3])
{
  double s2s[3], ray[3], m2s[3], a[3], b[3], tmp[3], sol[3];
  for (int i = 0; i < 3; ++i)
  {
    s2s[i] = sample[i] - sensor[i];
  }
  const double inv_len = 1.0 / sqrt(dot3(s2s, s2s)); /* glm::normalize */
  for (int i = 0; i < 3; ++i)
  {
    ray[i] = s2s[i] * inv_len;
    m2s[i] = sensor[i] - mean[i];
  }
  solve_triangular(cov, ray, a);
  solve_triangular(cov, m2s, b);
  const double t = -dot3(a, b) / dot3(a, a);
  for (int i = 0; i < 3; ++i)
  {
    x_ml[i] = ray[i] * t + sensor[i];
    tmp[i] = x_ml[i] - mean[i];
  }
  solve_triangular(cov, tmp, sol);
  *p_voxel = exp(-0.5 * dot3(sol, sol));
  const double noise_var = sensor_noise * sensor_noise; /* float*float then widened (:261) */
  for (int i = 0; i < 3; ++i)
  {
    tmp[i] = x_ml[i] - sample[i];
  }
  *p_sample = exp(-0.5 * dot3(tmp, tmp) / noise_var);
}

/* :542-635 calculateMissNdt */
void oracle_calculate_miss_ndt(const float cov[6], float *value, int *is_miss, const double sensor[3],
                               const double sample[3], const double mean[3], uint32_t count, float uninit,
                               float miss_value, float adaptation_rate, float sensor_noise, uint32_t sample_threshold)
{
  if (*value == uninit)
  {
    *value = miss_value;
    *is_miss = 1;
    return;
  }
  if (count < sample_threshold)
  {
    *value += miss_value;
    *is_miss = 1;
    return;
  }
  double p_voxel, p_sample, x_ml[3];
  sample_likelihoods(cov, sensor, sample, mean, sensor_noise, &p_voxel, &p_sample, x_ml);
  const double scaling = 0.5 * adaptation_rate;
  const double prod = p_voxel * (1.0 - p_sample);
  const double update = 0.5 - scaling * prod;
  *is_miss = prod < scaling;
  if (update == update)
  {
    *value += (float)log(update / (1.0 - update));
  }
}

/* :301-375 calculateHitWithCovariance (with :90-98 initialiseCovariance, :146-163 unpackCovariance) */
int oracle_calculate_hit_with_covariance(float cov[6], float *value, const double sample[3], const double mean[3],
                                         uint32_t count, float hit_value, float uninit, float resolution,
                                         float reinit_threshold, uint32_t reinit_count)
{
  const float initial = *value;
  const int was_uncertain = initial == uninit;
  int initialised = 0;
  if (count == 0 || (initial < reinit_threshold && count >= reinit_count))
  {
    cov[0] = cov[2] = cov[5] = 0.1f * resolution;
    cov[1] = cov[3] = cov[4] = 0;
    initialised = 1;
    count = 0;
  }
  *value = (!was_uncertain) ? hit_value + initial : hit_value;

  double s2m[3];
  for (int i = 0; i < 3; ++i)
  {
    s2m[i] = (!initialised) ? sample[i] - mean[i] : 0.0;
  }
  double A[9];
  const double one_on = (double)1 / (count + (double)1);
  const double sc_1 = count ? sqrt(count * one_on) : (double)1;
  const double sc_2 = one_on * sqrt((double)count);
  for (int i = 0; i < 6; ++i)
  {
    A[i] = sc_1 * cov[i];
  }
  A[6] = sc_2 * s2m[0];
  A[7] = sc_2 * s2m[1];
  A[8] = sc_2 * s2m[2];

  for (int k = 0; k < 3; ++k)
  {
    const int ind1 = (k * (k + 3)) >> 1;
    const int indk = ind1 - k;
    const double ak = sqrt(packed_dot(A, k, k));
    cov[ind1] = (float)ak;
    if (ak > 0)
    {
      const double aki = (double)1 / ak;
      for (int j = k + 1; j < 3; ++j)
      {
        const int indj = (j * (j + 1)) >> 1;
        const int indkj = indj + k;
        double c = packed_dot(A, j, k) * aki;
        cov[indkj] = (float)c;
        c *= aki;
        A[j + 6] -= c * A[k + 6];
        for (int l = 0; l <= k; ++l)
        {
          A[indj + l] -= c * A[indk + l];
        }
      }
    }
  }
  return initialised;
}

/* :391-411 calculateIntensityUpdateOnHit */
static void intensity_update_on_hit(float im[2], float value, float sample, float initial_cov, uint32_t count,
                                    float reinit_threshold, uint32_t reinit_count)
{
  const int needs_reset = count == 0 || (value < reinit_threshold && count >= reinit_count);
  const float delta = im[0] - sample;
  const float n = (float)count;
  const float inv = 1.0f / (n + 1.0f);
  const float mean = (!needs_reset) ? inv * (n * im[0] + sample) : sample;
  const float cv = (!needs_reset) ? inv * (n * im[1] + inv * delta * delta) : initial_cov;
  im[0] = mean;
  im[1] = cv;
}

/* :447-505 calculateHitMissUpdateOnHit */
static void hit_miss_update_on_hit(const float cov[6], float value, uint32_t hm[2], const double sensor[3],
                                   const double sample[3], const double mean[3], uint32_t count, float uninit,
                                   int reinit_perm, float adaptation_rate, float sensor_noise, float reinit_threshold,
                                   uint32_t reinit_count, uint32_t sample_threshold)
{
  const int needs_reset =
    value == uninit || (reinit_perm && (count == 0 || (value < reinit_threshold && count >= reinit_count)));
  const uint32_t initial_hit = (!needs_reset) ? hm[0] : 0;
  const uint32_t initial_miss = (!needs_reset) ? hm[1] : 0;
  double p_voxel, p_sample, x_ml[3];
  sample_likelihoods(cov, sensor, sample, mean, sensor_noise, &p_voxel, &p_sample, x_ml);
  const double prod = p_voxel * p_sample;
  const double eta = 0.5 * adaptation_rate;
  const int inc_hit = needs_reset || count < sample_threshold || (count >= sample_threshold && prod >= eta);
  const int inc_miss = !needs_reset && count >= sample_threshold && prod < eta && p_voxel >= eta;
  hm[0] = initial_hit + (inc_hit ? 1u : 0u);
  hm[1] = initial_miss + (inc_miss ? 1u : 0u);
}

/* ------------------------------------------------------------------------------------------ */
/* TSDF: ohm/VoxelTsdfCompute.h:57-136                                                         */
/* ------------------------------------------------------------------------------------------ */
int oracle_calculate_tsdf(const double sensor[3], const double sample[3], const double centre[3], float trunc,
                          float max_weight, float dropoff, float sparsity, float *weight, float *distance)
{
  double s2v[3], s2s[3];
  for (int i = 0; i < 3; ++i)
  {
    s2v[i] = centre[i] - sensor[i];
    s2s[i] = sample[i] - sensor[i];
  }
  const float distance_g = (float)sqrt(dot3(s2s, s2s));
  const float distance_g_v = (float)dot3(s2v, s2s) / distance_g;
  const float sdf = distance_g - distance_g_v;

  const float initial_weight = *weight;
  float updated_weight = 1.0f;
  updated_weight *= (dropoff > 0) ? ((trunc + sdf) / (trunc - dropoff)) : 1.0f;
  updated_weight = fmaxf(updated_weight, 0.0f);
  updated_weight *= (sparsity > 0 && fabsf(sdf) < trunc) ? sparsity : 1.0f;
  const float new_weight = initial_weight + updated_weight;
  const float abs_new_weight = fabsf(new_weight);
  const int near_zero = abs_new_weight < 0.00001f;
  const float new_sdf = (!near_zero) ? (sdf * updated_weight + *distance * initial_weight) / new_weight : 0.0f;
  *distance = (!near_zero) ? ((new_sdf > 0.0f) ? fminf(trunc, new_sdf) : fmaxf(-trunc, new_sdf)) : *distance;
  *weight = (!near_zero) ? fminf(new_weight, max_weight) : initial_weight;
  return !near_zero;
}

/* ------------------------------------------------------------------------------------------ */
/* Mappers                                                                                     */
/* ------------------------------------------------------------------------------------------ */
#define RF_END_POINT_AS_FREE (1u << 0)
#define RF_STOP_ON_FIRST_OCCUPIED (1u << 1)
#define RF_EXCLUDE_ORIGIN (1u << 2)
#define RF_EXCLUDE_SAMPLE (1u << 3)
#define RF_EXCLUDE_RAY (1u << 4)
#define RF_EXCLUDE_UNOBSERVED (1u << 5)
#define RF_EXCLUDE_FREE (1u << 6)
#define RF_EXCLUDE_OCCUPIED (1u << 7)

static size_t voxel_index(const oracle_map *m, const int32_t key[6])
{
  /* ohm/MapChunk.h:47-50 */
  return (size_t)key[3] + (size_t)key[4] * m->p.region_dim[0] +
         (size_t)key[5] * m->p.region_dim[0] * m->p.region_dim[1];
}

static chunk *key_chunk(oracle_map *m, const int32_t key[6])
{
  const int16_t r[3] = { (int16_t)key[0], (int16_t)key[1], (int16_t)key[2] };
  return map_region(m, r, 1);
}

typedef struct occ_ctx
{
  oracle_map *m;
  unsigned ray_flags;
  float sat_min, sat_max;
  int stop_adjustments;
  double last_exit_range;
  /* NDT */
  const double *sensor;
  const double *sample;
} occ_ctx;

/* RayMapperOccupancy.cpp:105-193 visit_func */
static int occupancy_visit(void *vctx, const int32_t key[6], double enter_range, double exit_range)
{
  occ_ctx *c = (occ_ctx *)vctx;
  oracle_map *m = c->m;
  chunk *ch = key_chunk(m, key);
  const size_t vi = voxel_index(m, key);
  float *occ = (float *)ch->layers[ORC_LAYER_OCCUPANCY];
  const float initial = occ[vi];
  const float uninit = INFINITY;
  const int unobserved = initial == uninit;
  const int is_free = !unobserved && initial < m->p.threshold_value;
  const int occupied = !unobserved && initial >= m->p.threshold_value;

  float adj = m->p.miss_value;
  adj = (unobserved && (c->ray_flags & RF_EXCLUDE_UNOBSERVED)) ? uninit : adj;
  adj = (is_free && (c->ray_flags & RF_EXCLUDE_FREE)) ? 0.0f : adj;
  adj = (occupied && (c->ray_flags & RF_EXCLUDE_OCCUPIED)) ? 0.0f : adj;

  float value;
  oracle_occupancy_adjust_miss(&value, initial, adj, uninit, m->p.min_value, c->sat_min, c->sat_max,
                               c->stop_adjustments);
  occ[vi] = value;

  if (ch->layers[ORC_LAYER_TRAVERSAL])
  {
    float *trav = (float *)ch->layers[ORC_LAYER_TRAVERSAL];
    trav[vi] += (float)(exit_range - enter_range);
  }
  c->stop_adjustments = c->stop_adjustments || ((c->ray_flags & RF_STOP_ON_FIRST_OCCUPIED) && occupied);
  c->last_exit_range = exit_range;
  ++m->stats.voxel_visits;
  return 1;
}

static double length3(const double a[3], const double b[3])
{
  const double d[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
  return sqrt(dot3(d, d));
}

/* RayMapperOccupancy.cpp:68-339 */
size_t oracle_integrate_occupancy(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                                  const double *timestamps, unsigned ray_flags)
{
  (void)intensities;
  occ_ctx c;
  memset(&c, 0, sizeof(c));
  c.m = m;
  c.ray_flags = ray_flags;
  c.sat_min = m->p.saturate_min ? m->p.min_value : -3.402823466e+38f;
  c.sat_max = m->p.saturate_max ? m->p.max_value : 3.402823466e+38f;
  const float uninit = INFINITY;

  if (timestamps && m->first_ray_time < 0)
  {
    m->first_ray_time = timestamps[0]; /* OccupancyMap::updateFirstRayTime */
  }
  const double time_base = m->first_ray_time;

  for (size_t i = 0; i < element_count; i += 2)
  {
    unsigned filter_flags = 0;
    double start[3] = { rays[3 * i], rays[3 * i + 1], rays[3 * i + 2] };
    double end[3] = { rays[3 * i + 3], rays[3 * i + 4], rays[3 * i + 5] };
    ++m->stats.rays_in;
    if (!apply_filter(m, start, end, &filter_flags))
    {
      continue;
    }
    ++m->stats.rays_accepted;

    const int include_sample = (filter_flags & RFF_CLIPPED_END) || (ray_flags & RF_END_POINT_AS_FREE);
    unsigned walk_flags = (!include_sample) ? ORC_WALK_EXCLUDE_END : 0u;
    walk_flags |= (ray_flags & RF_EXCLUDE_ORIGIN) ? ORC_WALK_EXCLUDE_START : 0u;

    if (!(ray_flags & RF_EXCLUDE_RAY))
    {
      c.stop_adjustments = 0;
      walk_segment_keys(m, start, end, walk_flags, occupancy_visit, &c);
    }

    if (!c.stop_adjustments && !include_sample && !(ray_flags & RF_EXCLUDE_SAMPLE))
    {
      int32_t key[6];
      if (!oracle_voxel_key(m, end, key))
      {
        continue; /* the reference would dereference a null key here; unreachable for finite input */
      }
      chunk *ch = key_chunk(m, key);
      const size_t vi = voxel_index(m, key);
      float *occ = (float *)ch->layers[ORC_LAYER_OCCUPANCY];
      const float initial = occ[vi];
      const int unobserved = initial == uninit;
      const int is_free = !unobserved && initial < m->p.threshold_value;
      const int occupied = !unobserved && initial >= m->p.threshold_value;
      float adj = m->p.hit_value;
      adj = (unobserved && (ray_flags & RF_EXCLUDE_UNOBSERVED)) ? uninit : adj;
      adj = (is_free && (ray_flags & RF_EXCLUDE_FREE)) ? 0.0f : adj;
      adj = (occupied && (ray_flags & RF_EXCLUDE_OCCUPIED)) ? 0.0f : adj;
      float value;
      oracle_occupancy_adjust_hit(&value, initial, adj, uninit, m->p.max_value, c.sat_min, c.sat_max,
                                  c.stop_adjustments);

      uint32_t sample_count = 0;
      if (ch->layers[ORC_LAYER_MEAN])
      {
        uint32_t *mean = (uint32_t *)ch->layers[ORC_LAYER_MEAN] + 2 * vi;
        double centre[3], local[3];
        oracle_voxel_centre(m, key, centre);
        for (int a = 0; a < 3; ++a)
        {
          local[a] = end[a] - centre[a];
        }
        mean[0] = oracle_sub_voxel_update(mean[0], mean[1], local, m->p.resolution);
        sample_count = mean[1];
        ++mean[1];
      }
      occ[vi] = value;

      if (ch->layers[ORC_LAYER_TRAVERSAL])
      {
        float *trav = (float *)ch->layers[ORC_LAYER_TRAVERSAL];
        trav[vi] += (float)(length3(end, start) - c.last_exit_range);
      }
      if (ch->layers[ORC_LAYER_TOUCH_TIME] && timestamps)
      {
        ((uint32_t *)ch->layers[ORC_LAYER_TOUCH_TIME])[vi] = oracle_encode_touch_time(time_base, timestamps[i >> 1]);
      }
      if (ch->layers[ORC_LAYER_INCIDENT])
      {
        uint32_t *inc = (uint32_t *)ch->layers[ORC_LAYER_INCIDENT];
        /* dvec3 -> glm::vec3 narrowing of (start - end) */
        const float ray[3] = { (float)(start[0] - end[0]), (float)(start[1] - end[1]), (float)(start[2] - end[2]) };
        inc[vi] = oracle_update_incident_normal(inc[vi], ray, sample_count);
      }
      ++m->stats.sample_updates;
    }
  }
  return element_count / 2;
}

/* RayMapperNdt.cpp:140-244 visit_func */
static int ndt_visit(void *vctx, const int32_t key[6], double enter_range, double exit_range)
{
  occ_ctx *c = (occ_ctx *)vctx;
  oracle_map *m = c->m;
  chunk *ch = key_chunk(m, key);
  const size_t vi = voxel_index(m, key);
  float *occ = (float *)ch->layers[ORC_LAYER_OCCUPANCY];
  const float *cov = (const float *)ch->layers[ORC_LAYER_COVARIANCE] + 6 * vi;
  const uint32_t *vmean = (const uint32_t *)ch->layers[ORC_LAYER_MEAN] + 2 * vi;
  double mean[3], centre[3];
  oracle_sub_voxel_to_local(vmean[0], m->p.resolution, mean);
  oracle_voxel_centre(m, key, centre);
  for (int a = 0; a < 3; ++a)
  {
    mean[a] += centre[a];
  }
  const float initial = occ[vi];
  float adjusted = initial;
  int is_miss = 0;
  oracle_calculate_miss_ndt(cov, &adjusted, &is_miss, c->sensor, c->sample, mean, vmean[1], INFINITY, m->p.miss_value,
                            m->p.adaptation_rate, m->p.sensor_noise, m->p.sample_threshold);
  if (m->p.ndt_tm && ch->layers[ORC_LAYER_HIT_MISS])
  {
    uint32_t *hm = (uint32_t *)ch->layers[ORC_LAYER_HIT_MISS] + 2 * vi;
    hm[1] += is_miss ? 1u : 0u;
  }
  float value;
  occupancy_adjust_down(&value, initial, adjusted, INFINITY, m->p.min_value, c->sat_min, c->sat_max,
                        c->stop_adjustments);
  occ[vi] = value;
  if (ch->layers[ORC_LAYER_TRAVERSAL])
  {
    float *trav = (float *)ch->layers[ORC_LAYER_TRAVERSAL];
    trav[vi] += (float)(exit_range - enter_range);
  }
  c->last_exit_range = exit_range;
  ++m->stats.voxel_visits;
  return 1;
}

/* RayMapperNdt.cpp:84-407 */
size_t oracle_integrate_ndt(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                            const double *timestamps, unsigned ray_flags)
{
  occ_ctx c;
  memset(&c, 0, sizeof(c));
  c.m = m;
  c.ray_flags = ray_flags;
  c.sat_min = m->p.saturate_min ? m->p.min_value : -3.402823466e+38f;
  c.sat_max = m->p.saturate_max ? m->p.max_value : 3.402823466e+38f;
  if (timestamps && m->first_ray_time < 0)
  {
    m->first_ray_time = timestamps[0];
  }
  const double time_base = m->first_ray_time;
  float intensity = 0.0f;

  for (size_t i = 0; i < element_count; i += 2)
  {
    unsigned filter_flags = 0;
    double start[3] = { rays[3 * i], rays[3 * i + 1], rays[3 * i + 2] };
    double sample[3] = { rays[3 * i + 3], rays[3 * i + 4], rays[3 * i + 5] };
    if (intensities)
    {
      intensity = intensities[i >> 1];
    }
    ++m->stats.rays_in;
    if (!apply_filter(m, start, sample, &filter_flags))
    {
      continue;
    }
    ++m->stats.rays_accepted;
    c.sensor = start;
    c.sample = sample;

    const int include_sample = (filter_flags & RFF_CLIPPED_END) || (ray_flags & RF_END_POINT_AS_FREE);
    unsigned walk_flags = (!include_sample) ? ORC_WALK_EXCLUDE_END : 0u;
    walk_flags |= (ray_flags & RF_EXCLUDE_ORIGIN) ? ORC_WALK_EXCLUDE_START : 0u;
    if (!(ray_flags & RF_EXCLUDE_RAY))
    {
      c.stop_adjustments = 0;
      walk_segment_keys(m, start, sample, walk_flags, ndt_visit, &c);
    }

    if (!c.stop_adjustments && !include_sample)
    {
      int32_t key[6];
      if (!oracle_voxel_key(m, sample, key))
      {
        continue;
      }
      chunk *ch = key_chunk(m, key);
      const size_t vi = voxel_index(m, key);
      float *occ = (float *)ch->layers[ORC_LAYER_OCCUPANCY];
      float *cov = (float *)ch->layers[ORC_LAYER_COVARIANCE] + 6 * vi;
      uint32_t *vmean = (uint32_t *)ch->layers[ORC_LAYER_MEAN] + 2 * vi;
      double centre[3], mean[3], local[3];
      oracle_voxel_centre(m, key, centre);
      oracle_sub_voxel_to_local(vmean[0], m->p.resolution, mean);
      for (int a = 0; a < 3; ++a)
      {
        mean[a] += centre[a];
      }
      const float initial = occ[vi];
      float adjusted = initial;

      if (m->p.ndt_tm && ch->layers[ORC_LAYER_INTENSITY] && ch->layers[ORC_LAYER_HIT_MISS])
      {
        float *im = (float *)ch->layers[ORC_LAYER_INTENSITY] + 2 * vi;
        uint32_t *hm = (uint32_t *)ch->layers[ORC_LAYER_HIT_MISS] + 2 * vi;
        hit_miss_update_on_hit(cov, adjusted, hm, start, sample, mean, vmean[1], INFINITY, 1, m->p.adaptation_rate,
                               m->p.sensor_noise, m->p.reinit_threshold, m->p.reinit_count, m->p.sample_threshold);
        intensity_update_on_hit(im, adjusted, intensity, m->p.initial_intensity_cov, vmean[1], m->p.reinit_threshold,
                                m->p.reinit_count);
      }

      const int reset_mean = oracle_calculate_hit_with_covariance(
        cov, &adjusted, sample, mean, vmean[1], m->p.hit_value, INFINITY, (float)m->p.resolution,
        m->p.reinit_threshold, m->p.reinit_count);
      float value;
      occupancy_adjust_up(&value, initial, adjusted, INFINITY, m->p.max_value, c.sat_min, c.sat_max,
                          c.stop_adjustments);
      vmean[1] = (!reset_mean) ? vmean[1] : 0;
      for (int a = 0; a < 3; ++a)
      {
        local[a] = sample[a] - centre[a];
      }
      vmean[0] = oracle_sub_voxel_update(vmean[0], vmean[1], local, m->p.resolution);
      ++vmean[1];
      occ[vi] = value;

      if (ch->layers[ORC_LAYER_TRAVERSAL])
      {
        float *trav = (float *)ch->layers[ORC_LAYER_TRAVERSAL];
        trav[vi] += (float)(length3(sample, start) - c.last_exit_range);
      }
      if (ch->layers[ORC_LAYER_TOUCH_TIME] && timestamps)
      {
        ((uint32_t *)ch->layers[ORC_LAYER_TOUCH_TIME])[vi] = oracle_encode_touch_time(time_base, timestamps[i >> 1]);
      }
      if (ch->layers[ORC_LAYER_INCIDENT])
      {
        uint32_t *inc = (uint32_t *)ch->layers[ORC_LAYER_INCIDENT];
        const float ray[3] = { (float)(start[0] - sample[0]), (float)(start[1] - sample[1]),
                               (float)(start[2] - sample[2]) };
        inc[vi] = oracle_update_incident_normal(inc[vi], ray, vmean[1] - 1);
      }
      ++m->stats.sample_updates;
    }
  }
  return element_count / 2;
}

typedef struct tsdf_ctx
{
  oracle_map *m;
  const double *sensor;
  const double *sample;
} tsdf_ctx;

/* RayMapperTsdf.cpp:114-160 visit_func */
static int tsdf_visit(void *vctx, const int32_t key[6], double enter_range, double exit_range)
{
  (void)enter_range;
  (void)exit_range;
  tsdf_ctx *c = (tsdf_ctx *)vctx;
  oracle_map *m = c->m;
  chunk *ch = key_chunk(m, key);
  const size_t vi = voxel_index(m, key);
  float *tsdf = (float *)ch->layers[ORC_LAYER_TSDF] + 2 * vi;
  double centre[3];
  oracle_voxel_centre(m, key, centre);
  oracle_calculate_tsdf(c->sensor, c->sample, centre, m->p.tsdf_trunc, m->p.tsdf_max_weight, m->p.tsdf_dropoff,
                        m->p.tsdf_sparsity, &tsdf[0], &tsdf[1]);
  ++m->stats.voxel_visits;
  return 1;
}

/* RayMapperTsdf.cpp:87-182 (ray flags ignored there) */
size_t oracle_integrate_tsdf(oracle_map *m, const double *rays, size_t element_count, const float *intensities,
                             const double *timestamps, unsigned ray_flags)
{
  (void)intensities;
  (void)ray_flags;
  if (timestamps && m->first_ray_time < 0)
  {
    m->first_ray_time = timestamps[0];
  }
  tsdf_ctx c;
  c.m = m;
  for (size_t i = 0; i < element_count; i += 2)
  {
    unsigned filter_flags = 0;
    const double *sensor = rays + 3 * i;
    const double *sample = rays + 3 * i + 3;
    double start[3] = { sensor[0], sensor[1], sensor[2] };
    double end[3] = { sample[0], sample[1], sample[2] };
    ++m->stats.rays_in;
    if (!apply_filter(m, start, end, &filter_flags))
    {
      continue;
    }
    ++m->stats.rays_accepted;
    c.sensor = sensor;
    c.sample = sample;
    walk_segment_keys(m, start, end, 0u, tsdf_visit, &c);
  }
  return element_count / 2;
}

/* ------------------------------------------------------------------------------------------ */
/* RaysQuery::onExecute (ohm/RaysQuery.cpp:109-199)                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct rays_query_ctx
{
  const oracle_map *m;
  double volume_coefficient;
  double unobserved_volume;
  float range;
  int terminal_state;
  int32_t terminal_key[6];
  chunk *last_chunk;
  int16_t last_region[3];
} rays_query_ctx;

/* the visit lambda, RaysQuery.cpp:129-158 */
static int rays_query_visit(void *vctx, const int32_t key[6], double enter_range, double exit_range)
{
  rays_query_ctx *c = (rays_query_ctx *)vctx;
  const oracle_map *m = c->m;
  const int16_t region[3] = { (int16_t)key[0], (int16_t)key[1], (int16_t)key[2] };
  const size_t voxel_index =
    (size_t)key[3] + (size_t)key[4] * m->p.region_dim[0] + (size_t)key[5] * m->p.region_dim[0] * m->p.region_dim[1];
  float occupancy_value = INFINITY; /* unobservedOccupancyValue() */
  chunk *ch = (c->last_chunk && region[0] == c->last_region[0] && region[1] == c->last_region[1] &&
               region[2] == c->last_region[2]) ?
                c->last_chunk :
                map_region((oracle_map *)m, region, 0);
  if (ch)
  {
    occupancy_value = ((const float *)ch->layers[ORC_LAYER_OCCUPANCY])[voxel_index];
    c->last_region[0] = region[0];
    c->last_region[1] = region[1];
    c->last_region[2] = region[2];
  }
  c->last_chunk = ch;
  const int is_unobserved = occupancy_value == INFINITY;
  const int is_occupied = !is_unobserved && occupancy_value > m->p.threshold_value;
  c->unobserved_volume +=
    is_unobserved ?
      (c->volume_coefficient * (exit_range * exit_range * exit_range - enter_range * enter_range * enter_range)) :
      0.0f;
  c->range = (!is_occupied) ? (float)exit_range : c->range;
  c->terminal_state = is_unobserved ? ORC_OCCUPANCY_UNOBSERVED : (is_occupied ? ORC_OCCUPANCY_OCCUPIED : ORC_OCCUPANCY_FREE);
  memcpy(c->terminal_key, key, sizeof(c->terminal_key));
  return !is_occupied;
}

size_t oracle_rays_query(const oracle_map *m, const double *rays, size_t element_count, double volume_coefficient,
                         double *ranges, double *unobserved_volumes, int *terminal_states, int32_t *terminal_keys)
{
  rays_query_ctx c;
  memset(&c, 0, sizeof(c));
  c.m = m;
  c.volume_coefficient = volume_coefficient;
  c.terminal_state = ORC_OCCUPANCY_NULL; /* RaysQuery.cpp:119-120: declared outside the ray loop, never reset */
  size_t n = 0;
  for (size_t i = 0; i + 1 < element_count; i += 2, ++n)
  {
    double start[3] = { rays[3 * i], rays[3 * i + 1], rays[3 * i + 2] };
    double end[3] = { rays[3 * i + 3], rays[3 * i + 4], rays[3 * i + 5] };
    unsigned filter_flags = 0;
    c.unobserved_volume = 0.0;
    c.range = 0.0f;
    if (!apply_filter(m, start, end, &filter_flags))
    {
      ranges[n] = c.range;
      unobserved_volumes[n] = c.unobserved_volume;
      terminal_states[n] = ORC_OCCUPANCY_NULL;
      for (int k = 0; k < 6; ++k)
      {
        terminal_keys[6 * n + k] = 0; /* Key::kNull is reported as all zero with *_states == NULL */
      }
      continue;
    }
    walk_segment_keys(m, start, end, 0u, rays_query_visit, &c);
    ranges[n] = c.range;
    unobserved_volumes[n] = c.unobserved_volume;
    terminal_states[n] = c.terminal_state;
    memcpy(terminal_keys + 6 * n, c.terminal_key, sizeof(c.terminal_key));
  }
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* RayMapperSecondarySample (ohm/RayMapperSecondarySample.cpp:37-74)                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct voxel_secondary_sample /* VoxelSecondarySample.h:29-38 */
{
  float m2;
  uint16_t range_mean;
  uint16_t count;
} voxel_secondary_sample;

/* addSecondarySample, VoxelSecondarySample.h:87-99 (Welford) */
static void add_secondary_sample(voxel_secondary_sample *voxel, double range)
{
  const double quantisation = 1000.0;                  /* secondarySampleQuantisationFactor() */
  const double max_range = (65535 - 1u) / quantisation; /* secondarySampleMaxRange() */
  range = (range < max_range) ? range : max_range;     /* std::min(range, max) */
  double range_mean = voxel->range_mean / quantisation;
  ++voxel->count;
  const double delta = range - range_mean;
  range_mean += delta / voxel->count;
  voxel->range_mean = (uint16_t)(range_mean * quantisation);
  const double delta2 = range - range_mean;
  voxel->m2 += (float)(delta * delta2);
}

size_t oracle_integrate_secondary(oracle_map *m, const double *rays, size_t element_count)
{
  if (!(m->p.layers & (1u << ORC_LAYER_SECONDARY)))
  {
    return 0;
  }
  for (size_t i = 0; i + 1 < element_count; i += 2)
  {
    const double *start = rays + 3 * i, *end = rays + 3 * i + 3;
    const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
    const double range = sqrt(dot3(d, d)); /* glm::length */
    int32_t key[6];
    if (!oracle_voxel_key(m, end, key))
    {
      continue;
    }
    const int16_t region[3] = { (int16_t)key[0], (int16_t)key[1], (int16_t)key[2] };
    chunk *ch = map_region(m, region, 1);
    const size_t voxel_index =
      (size_t)key[3] + (size_t)key[4] * m->p.region_dim[0] + (size_t)key[5] * m->p.region_dim[0] * m->p.region_dim[1];
    add_secondary_sample((voxel_secondary_sample *)ch->layers[ORC_LAYER_SECONDARY] + voxel_index, range);
    ++m->stats.sample_updates;
  }
  return element_count / 2;
}
