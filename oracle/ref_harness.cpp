// ref_harness.cpp — C entry points over the REFERENCE's own CPU mappers (ohm::RayMapperOccupancy / RayMapperNdt /
// RayMapperTsdf on an ohm::OccupancyMap), compiled unmodified from /root/reference into oracle/_ref/libohm_ref.so
// by oracle/Makefile (`make ref`).  TEST INFRASTRUCTURE ONLY: it validates the C restatement (oracle/ohm_oracle.c)
// against the real thing and serves as the "reference" CPU baseline of bench.py.  Nothing here is reference code;
// it only calls the reference's public API.
#include "ohm_oracle.h"

#include <ohm/Aabb.h>
#include <ohm/DefaultLayer.h>
#include <ohm/MapChunk.h>
#include <ohm/MapLayer.h>
#include <ohm/MapLayout.h>
#include <ohm/NdtMap.h>
#include <ohm/OccupancyMap.h>
#include <ohm/RayFilter.h>
#include <ohm/RayMapperNdt.h>
#include <ohm/RayMapperOccupancy.h>
#include <ohm/RayMapperTsdf.h>
#include <ohm/RayPattern.h>
#include <ohm/VoxelBuffer.h>
#include <ohm/VoxelTsdf.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

// The two RayPattern members referenced (never called) by translation units on the path; RayPattern.cpp needs
// quaternion/matrix products the GLM stand-in does not provide.
size_t ohm::RayPattern::buildRays(std::vector<glm::dvec3> *, const glm::dmat4 &) const
{
  abort();
}
size_t ohm::RayPattern::buildRays(std::vector<glm::dvec3> *, const glm::dvec3 &, const glm::dquat &, double) const
{
  abort();
}

namespace
{
struct RefMap
{
  std::unique_ptr<ohm::OccupancyMap> map;
  std::unique_ptr<ohm::NdtMap> ndt;
  std::unique_ptr<ohm::RayMapper> mapper;
  int mode = 0;
};

const char *layerName(int layer)
{
  using namespace ohm::default_layer;
  switch (layer)
  {
  case ORC_LAYER_OCCUPANCY: return occupancyLayerName();
  case ORC_LAYER_MEAN: return meanLayerName();
  case ORC_LAYER_TRAVERSAL: return traversalLayerName();
  case ORC_LAYER_TOUCH_TIME: return touchTimeLayerName();
  case ORC_LAYER_INCIDENT: return incidentNormalLayerName();
  case ORC_LAYER_COVARIANCE: return covarianceLayerName();
  case ORC_LAYER_SECONDARY: return secondarySamplesLayerName();
  case ORC_LAYER_INTENSITY: return intensityLayerName();
  case ORC_LAYER_HIT_MISS: return hitMissCountLayerName();
  case ORC_LAYER_TSDF: return tsdfLayerName();
  default: return "";
  }
}

void applyParams(RefMap &r, const oracle_params &p)
{
  ohm::OccupancyMap &m = *r.map;
  m.setOrigin(glm::dvec3(p.origin[0], p.origin[1], p.origin[2]));
  m.setHitValue(p.hit_value);
  m.setMissValue(p.miss_value);
  m.setMinVoxelValue(p.min_value);
  m.setMaxVoxelValue(p.max_value);
  m.setSaturateAtMinValue(p.saturate_min != 0);
  m.setSaturateAtMaxValue(p.saturate_max != 0);
  const double range = p.filter_range;
  switch (p.filter_kind)
  {
  case ORC_FILTER_NONE:
    m.setRayFilter(ohm::RayFilterFunction());
    break;
  case ORC_FILTER_CLIP_BOX: {
    const ohm::Aabb box(glm::dvec3(p.clip_box[0], p.clip_box[1], p.clip_box[2]),
                        glm::dvec3(p.clip_box[3], p.clip_box[4], p.clip_box[5]));
    m.setRayFilter([box](glm::dvec3 *s, glm::dvec3 *e, unsigned *f) { return ohm::clipBounded(s, e, f, box); });
    break;
  }
  case ORC_FILTER_CLIP_RANGE:
    m.setRayFilter([range](glm::dvec3 *s, glm::dvec3 *e, unsigned *f) { return ohm::clipRayFilter(s, e, f, range); });
    break;
  default:
    m.setRayFilter([range](glm::dvec3 *s, glm::dvec3 *e, unsigned *f) { return ohm::goodRayFilter(s, e, f, range); });
    break;
  }
  if (r.ndt)
  {
    r.ndt->setSensorNoise(p.sensor_noise);
    r.ndt->setAdaptationRate(p.adaptation_rate);
    r.ndt->setReinitialiseCovarianceThreshold(p.reinit_threshold);
    r.ndt->setReinitialiseCovariancePointCount(p.reinit_count);
    r.ndt->setNdtSampleThreshold(p.sample_threshold);
    r.ndt->setInitialIntensityCovariance(p.initial_intensity_cov);
  }
}
}  // namespace

extern "C" {

// mode: 0 occupancy, 1 ndt (NdtMode::kOccupancy), 2 ndt-tm, 3 tsdf
void *ref_map_create(const oracle_params *p, int mode)
{
  auto *r = new RefMap;
  r->mode = mode;
  ohm::MapFlag flags = ohm::MapFlag::kNone;
  if (p->layers & (1u << ORC_LAYER_MEAN)) flags |= ohm::MapFlag::kVoxelMean;
  if (p->layers & (1u << ORC_LAYER_TRAVERSAL)) flags |= ohm::MapFlag::kTraversal;
  if (p->layers & (1u << ORC_LAYER_TOUCH_TIME)) flags |= ohm::MapFlag::kTouchTime;
  if (p->layers & (1u << ORC_LAYER_INCIDENT)) flags |= ohm::MapFlag::kIncidentNormal;
  if (p->layers & (1u << ORC_LAYER_SECONDARY)) flags |= ohm::MapFlag::kSecondarySample;
  const glm::u8vec3 dim(uint8_t(p->region_dim[0]), uint8_t(p->region_dim[1]), uint8_t(p->region_dim[2]));
  r->map.reset(new ohm::OccupancyMap(p->resolution, dim, flags));
  if (mode == 3)
  {
    // ohmapp/OhmAppGpu.cpp:192-201: a layout holding only the TSDF layer
    ohm::MapLayout layout;
    ohm::addTsdf(layout);
    r->map->updateLayout(layout);
    auto *tsdf = new ohm::RayMapperTsdf(r->map.get());
    ohm::TsdfOptions o;
    o.max_weight = p->tsdf_max_weight;
    o.default_truncation_distance = p->tsdf_trunc;
    o.dropoff_epsilon = p->tsdf_dropoff;
    o.sparsity_compensation_factor = p->tsdf_sparsity;
    tsdf->setTsdfOptions(o);
    r->mapper.reset(tsdf);
  }
  else if (mode == 1 || mode == 2)
  {
    r->ndt.reset(new ohm::NdtMap(r->map.get(), true, mode == 2 ? ohm::NdtMode::kTraversability : ohm::NdtMode::kOccupancy));
    r->mapper.reset(new ohm::RayMapperNdt(r->ndt.get()));
  }
  else
  {
    r->mapper.reset(new ohm::RayMapperOccupancy(r->map.get()));
  }
  applyParams(*r, *p);
  return r;
}

void ref_map_destroy(void *h)
{
  auto *r = static_cast<RefMap *>(h);
  if (r)
  {
    r->mapper.reset();
    r->ndt.reset();
    r->map.reset();
    delete r;
  }
}

void ref_map_set_params(void *h, const oracle_params *p)
{
  applyParams(*static_cast<RefMap *>(h), *p);
}

int ref_mapper_valid(void *h)
{
  return static_cast<RefMap *>(h)->mapper->valid() ? 1 : 0;
}

size_t ref_integrate(void *h, const double *rays, size_t element_count, const float *intensities,
                     const double *timestamps, unsigned ray_flags)
{
  auto *r = static_cast<RefMap *>(h);
  static_assert(sizeof(glm::dvec3) == 3 * sizeof(double), "dvec3 layout");
  return r->mapper->integrateRays(reinterpret_cast<const glm::dvec3 *>(rays), element_count, intensities, timestamps,
                                  ray_flags);
}

double ref_first_ray_time(void *h)
{
  return static_cast<RefMap *>(h)->map->firstRayTime();
}

size_t ref_region_count(void *h)
{
  return static_cast<RefMap *>(h)->map->regionCount();
}

size_t ref_region_keys(void *h, int16_t *keys, size_t cap)
{
  std::vector<const ohm::MapChunk *> chunks;
  static_cast<RefMap *>(h)->map->enumerateRegions(chunks);
  std::vector<glm::i16vec3> all;
  for (const ohm::MapChunk *c : chunks)
  {
    all.push_back(c->region.coord);
  }
  std::sort(all.begin(), all.end(), [](const glm::i16vec3 &a, const glm::i16vec3 &b) {
    if (a.z != b.z) return a.z < b.z;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
  });
  for (size_t i = 0; i < all.size() && i < cap; ++i)
  {
    keys[3 * i] = all[i].x;
    keys[3 * i + 1] = all[i].y;
    keys[3 * i + 2] = all[i].z;
  }
  return all.size();
}

// Copies one region's layer block out.  Returns the number of bytes copied, 0 when the region/layer is absent.
size_t ref_region_layer(void *h, const int16_t key[3], int layer, void *dst, size_t bytes)
{
  auto *r = static_cast<RefMap *>(h);
  const int index = r->map->layout().layerIndex(layerName(layer));
  if (index < 0)
  {
    return 0;
  }
  ohm::MapChunk *chunk = r->map->region(glm::i16vec3(key[0], key[1], key[2]), false);
  if (!chunk)
  {
    return 0;
  }
  ohm::VoxelBuffer<const ohm::VoxelBlock> buffer(chunk->voxel_blocks[index]);
  const size_t n = std::min(bytes, buffer.voxelMemorySize());
  memcpy(dst, buffer.voxelMemory(), n);
  return n;
}

// Voxel keys along a segment, exactly as the mappers walk it (ohm/LineWalk.h:112-129).
size_t ref_walk_segment(void *h, const double start[3], const double end[3], unsigned walk_flags, int32_t *keys,
                        double *enter, double *exit, size_t cap);
}

#include <ohm/LineWalk.h>

extern "C" size_t ref_walk_segment(void *h, const double start[3], const double end[3], unsigned walk_flags, int32_t *keys,
                                   double *enter, double *exit, size_t cap)
{
  auto *r = static_cast<RefMap *>(h);
  size_t n = 0;
  const auto visit = [&](const ohm::Key &key, double enter_range, double exit_range) -> bool {
    if (n < cap)
    {
      keys[6 * n + 0] = key.regionKey().x;
      keys[6 * n + 1] = key.regionKey().y;
      keys[6 * n + 2] = key.regionKey().z;
      keys[6 * n + 3] = key.localKey().x;
      keys[6 * n + 4] = key.localKey().y;
      keys[6 * n + 5] = key.localKey().z;
      enter[n] = enter_range;
      exit[n] = exit_range;
    }
    ++n;
    return true;
  };
  ohm::walkSegmentKeys(ohm::LineWalkContext(*r->map, visit), glm::dvec3(start[0], start[1], start[2]),
                       glm::dvec3(end[0], end[1], end[2]), walk_flags);
  return n;
}

#include <ohm/RaysQuery.h>

// ohm::RaysQuery on the reference map (ohm/RaysQuery.cpp:109-199).  Null keys are reported as six zeros.
extern "C" size_t ref_rays_query(void *h, const double *rays, size_t element_count, double volume_coefficient,
                                 double *ranges, double *unobserved_volumes, int *terminal_states, int32_t *terminal_keys)
{
  auto *r = static_cast<RefMap *>(h);
  ohm::RaysQuery query;
  query.setMap(r->map.get());
  query.setVolumeCoefficient(volume_coefficient);
  query.setRays(reinterpret_cast<const glm::dvec3 *>(rays), element_count);
  if (!query.execute())
  {
    return 0;
  }
  const size_t n = query.numberOfResults();
  const double *q_ranges = query.ranges();  // stored as double, assigned from a float (RaysQuery.cpp:121,151)
  const double *q_volumes = query.unobservedVolumes();
  const ohm::OccupancyType *q_states = query.terminalOccupancyTypes();
  const ohm::Key *q_keys = query.intersectedVoxels();
  for (size_t i = 0; i < n; ++i)
  {
    ranges[i] = q_ranges[i];
    unobserved_volumes[i] = q_volumes[i];
    terminal_states[i] = static_cast<int>(q_states[i]);
    const bool null_key = q_keys[i].isNull();
    terminal_keys[6 * i + 0] = null_key ? 0 : q_keys[i].regionKey().x;
    terminal_keys[6 * i + 1] = null_key ? 0 : q_keys[i].regionKey().y;
    terminal_keys[6 * i + 2] = null_key ? 0 : q_keys[i].regionKey().z;
    terminal_keys[6 * i + 3] = null_key ? 0 : q_keys[i].localKey().x;
    terminal_keys[6 * i + 4] = null_key ? 0 : q_keys[i].localKey().y;
    terminal_keys[6 * i + 5] = null_key ? 0 : q_keys[i].localKey().z;
  }
  return n;
}

#include <ohm/LineKeysQuery.h>

// ohm::LineKeysQuery on the reference map (ohm/LineKeysQuery.cpp:103-123).  keys receives six int32 per key; returns the
// total number of keys (which may exceed key_capacity: nothing beyond the capacity is written).
extern "C" size_t ref_line_keys_query(void *h, const double *rays, size_t element_count, uint64_t *result_indices,
                                      uint64_t *result_counts, int32_t *keys, size_t key_capacity)
{
  auto *r = static_cast<RefMap *>(h);
  ohm::LineKeysQuery query(*r->map);
  query.setRays(reinterpret_cast<const glm::dvec3 *>(rays), element_count);
  if (!query.execute())
  {
    return 0;
  }
  const size_t n = query.numberOfResults();
  size_t total = 0;
  for (size_t i = 0; i < n; ++i)
  {
    result_indices[i] = query.resultIndices()[i];
    result_counts[i] = query.resultCounts()[i];
    total = std::max<size_t>(total, query.resultIndices()[i] + query.resultCounts()[i]);
  }
  const ohm::Key *q_keys = query.intersectedVoxels();
  for (size_t k = 0; k < total && k < key_capacity; ++k)
  {
    keys[6 * k + 0] = q_keys[k].regionKey().x;
    keys[6 * k + 1] = q_keys[k].regionKey().y;
    keys[6 * k + 2] = q_keys[k].regionKey().z;
    keys[6 * k + 3] = q_keys[k].localKey().x;
    keys[6 * k + 4] = q_keys[k].localKey().y;
    keys[6 * k + 5] = q_keys[k].localKey().z;
  }
  return total;
}

#include <ohm/MapSerialise.h>

// ohm::save / ohm::load (ohm/MapSerialise.cpp:595-705) on the reference map.  ref_load returns a new handle holding
// the loaded map (no mapper: for inspection with ref_region_keys / ref_region_layer only), or null.
extern "C" int ref_save(void *h, const char *path)
{
  auto *r = static_cast<RefMap *>(h);
  return ohm::save(path, *r->map);
}

extern "C" void *ref_load(const char *path, int *error)
{
  auto *r = new RefMap;
  r->map.reset(new ohm::OccupancyMap(1.0));
  const int err = ohm::load(path, *r->map);
  if (error)
  {
    *error = err;
  }
  if (err)
  {
    delete r;
    return nullptr;
  }
  return r;
}

extern "C" void ref_map_header(void *h, double *resolution, double origin[3], int region_dim[3], double *first_ray_time,
                               double *threshold, double *hit, double *miss, unsigned *flags)
{
  auto *r = static_cast<RefMap *>(h);
  *resolution = r->map->resolution();
  const glm::dvec3 o = r->map->origin();
  origin[0] = o.x;
  origin[1] = o.y;
  origin[2] = o.z;
  const glm::u8vec3 d = r->map->regionVoxelDimensions();
  region_dim[0] = d.x;
  region_dim[1] = d.y;
  region_dim[2] = d.z;
  *first_ray_time = r->map->firstRayTime();
  *threshold = r->map->occupancyThresholdValue();
  *hit = r->map->hitValue();
  *miss = r->map->missValue();
  *flags = unsigned(r->map->flags());
}

#include <ohm/RayMapperSecondarySample.h>

// ohm::RayMapperSecondarySample on the same reference map (ohm/RayMapperSecondarySample.cpp:37-74).
extern "C" size_t ref_integrate_secondary(void *h, const double *rays, size_t element_count)
{
  auto *r = static_cast<RefMap *>(h);
  ohm::RayMapperSecondarySample mapper(r->map.get());
  if (!mapper.valid())
  {
    return 0;
  }
  return mapper.integrateRays(reinterpret_cast<const glm::dvec3 *>(rays), element_count, nullptr, nullptr, 0u);
}
