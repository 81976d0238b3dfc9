// binding_harness.cpp — C entry points for tests/test_gpu_binding.py: the same rays through the binding (ohm::GpuMap /
// GpuNdtMap / GpuTsdfMap from GpuMapB200.cpp, i.e. libohmb200 behind ohm's own headers, on a real ohm::OccupancyMap) and
// through ohm's CPU mappers on a second map; then every MapChunk layer of the two host maps is compared.
// Test infrastructure; calls only public ohm API.
#include "GpuCacheB200.h"

#include <ohmgpu/GpuMap.h>
#include <ohmgpu/GpuNdtMap.h>
#include <ohmgpu/GpuTsdfMap.h>

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
#include <ohm/VoxelBuffer.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace
{
struct Side
{
  std::unique_ptr<ohm::OccupancyMap> map;
  std::unique_ptr<ohm::NdtMap> ndt;       // CPU side only (the GPU mappers own theirs)
  std::unique_ptr<ohm::RayMapper> mapper;
};

std::unique_ptr<ohm::OccupancyMap> makeMap(int mode, unsigned map_flags, double resolution)
{
  std::unique_ptr<ohm::OccupancyMap> map(new ohm::OccupancyMap(resolution, ohm::MapFlag(map_flags)));
  if (mode == 3)
  {
    ohm::MapLayout layout;  // ohmapp/OhmAppGpu.cpp:192-201: a layout holding only the TSDF layer
    ohm::addTsdf(layout);
    map->updateLayout(layout);
  }
  return map;
}

Side cpuSide(int mode, unsigned map_flags, double resolution)
{
  Side s;
  s.map = makeMap(mode, map_flags, resolution);
  if (mode == 1 || mode == 2)
  {
    s.ndt.reset(new ohm::NdtMap(s.map.get(), true, mode == 2 ? ohm::NdtMode::kTraversability : ohm::NdtMode::kOccupancy));
    s.mapper.reset(new ohm::RayMapperNdt(s.ndt.get()));
  }
  else if (mode == 3)
  {
    s.mapper.reset(new ohm::RayMapperTsdf(s.map.get()));
  }
  else
  {
    s.mapper.reset(new ohm::RayMapperOccupancy(s.map.get()));
  }
  return s;
}

Side gpuSide(int mode, unsigned map_flags, double resolution, size_t gpu_mem)
{
  Side s;
  s.map = makeMap(mode, map_flags, resolution);
  // the constructors of ohmgpu/GpuMap.h:167, GpuNdtMap.h:72, GpuTsdfMap.h:46, as tests/ohmtestgpu/GpuMapTest.cpp:90-92
  if (mode == 1 || mode == 2)
  {
    s.mapper.reset(new ohm::GpuNdtMap(s.map.get(), true, 2048u, gpu_mem,
                                      mode == 2 ? ohm::NdtMode::kTraversability : ohm::NdtMode::kOccupancy));
  }
  else if (mode == 3)
  {
    s.mapper.reset(new ohm::GpuTsdfMap(s.map.get(), true, 2048u, gpu_mem));
  }
  else
  {
    s.mapper.reset(new ohm::GpuMap(s.map.get(), true, 2048u, gpu_mem));
  }
  return s;
}

void integrate(ohm::RayMapper &mapper, const double *rays, size_t element_count, const float *intensities,
               const double *timestamps, unsigned ray_flags, size_t batch_rays)
{
  const auto *points = reinterpret_cast<const glm::dvec3 *>(rays);
  const size_t n = element_count / 2;
  const size_t step = batch_rays ? batch_rays : n;
  for (size_t first = 0; first < n; first += step)
  {
    const size_t count = std::min(step, n - first);
    mapper.integrateRays(points + 2 * first, 2 * count, intensities ? intensities + first : nullptr,
                         timestamps ? timestamps + first : nullptr, ray_flags);
  }
}

// Every layer of every chunk of `a` against `b`.  Returns the number of differing voxel words (-1: region sets differ).
// log_odds_tol > 0: the occupancy layer is compared as |a - b| <= tol (1 + |b|) (NDT), everything else bit for bit.
long long compareMaps(const ohm::OccupancyMap &a, const ohm::OccupancyMap &b, double log_odds_tol, std::string &why)
{
  std::vector<const ohm::MapChunk *> ca, cb;
  a.enumerateRegions(ca);
  b.enumerateRegions(cb);
  auto less = [](const ohm::MapChunk *x, const ohm::MapChunk *y) {
    const glm::i16vec3 p = x->region.coord, q = y->region.coord;
    return p.z != q.z ? p.z < q.z : (p.y != q.y ? p.y < q.y : p.x < q.x);
  };
  std::sort(ca.begin(), ca.end(), less);
  std::sort(cb.begin(), cb.end(), less);
  if (ca.size() != cb.size())
  {
    why = "region counts differ: " + std::to_string(ca.size()) + " vs " + std::to_string(cb.size());
    return -1;
  }
  if (a.layout().layerCount() != b.layout().layerCount())
  {
    why = "layer counts differ";
    return -1;
  }
  long long bad = 0;
  const int occupancy_layer = a.layout().occupancyLayer();
  for (size_t i = 0; i < ca.size(); ++i)
  {
    if (ca[i]->region.coord != cb[i]->region.coord)
    {
      why = "region sets differ";
      return -1;
    }
    for (size_t layer = 0; layer < a.layout().layerCount(); ++layer)
    {
      if (std::string(a.layout().layer(layer).name()) != b.layout().layer(layer).name())
      {
        why = "layer names differ";
        return -1;
      }
      ohm::VoxelBuffer<const ohm::VoxelBlock> va(ca[i]->voxel_blocks[layer]), vb(cb[i]->voxel_blocks[layer]);
      if (va.voxelMemorySize() != vb.voxelMemorySize())
      {
        why = "layer sizes differ";
        return -1;
      }
      const size_t words = va.voxelMemorySize() / 4;
      const auto *wa = reinterpret_cast<const uint32_t *>(va.voxelMemory());
      const auto *wb = reinterpret_cast<const uint32_t *>(vb.voxelMemory());
      for (size_t w = 0; w < words; ++w)
      {
        if (wa[w] == wb[w])
        {
          continue;
        }
        if (int(layer) == occupancy_layer && log_odds_tol > 0)
        {
          float fa, fb;
          memcpy(&fa, wa + w, 4);
          memcpy(&fb, wb + w, 4);
          if (std::fabs(double(fa) - double(fb)) <= log_odds_tol * (1.0 + std::fabs(double(fb))))
          {
            continue;
          }
        }
        if (bad == 0)
        {
          why = std::string("layer ") + a.layout().layer(layer).name() + " differs";
        }
        ++bad;
      }
    }
    if (ca[i]->first_valid_index != cb[i]->first_valid_index && occupancy_layer >= 0)
    {
      if (bad == 0)
      {
        why = "first_valid_index differs";
      }
      ++bad;
    }
  }
  return bad;
}

thread_local std::string g_why;
}  // namespace

extern "C" {

const char *binding_last_message()
{
  return g_why.c_str();
}

// The same rays through ohm::GpuMap/GpuNdtMap/GpuTsdfMap (mode 0 / 1,2 / 3) -> syncVoxels() and through the CPU mapper;
// returns the number of differing voxel words (0 = identical host maps), negative on a structural difference.
long long binding_compare(int mode, unsigned map_flags, double resolution, const double *rays, size_t element_count,
                          const float *intensities, const double *timestamps, unsigned ray_flags, size_t batch_rays,
                          double log_odds_tol, size_t gpu_mem, unsigned long long *regions_out)
{
  g_why.clear();
  Side gpu = gpuSide(mode, map_flags, resolution, gpu_mem);
  Side cpu = cpuSide(mode, map_flags, resolution);
  if (!gpu.mapper->valid())
  {
    g_why = std::string("GpuMap::gpuOk() is false: ") + ohmb200_last_error();
    return -2;
  }
  integrate(*gpu.mapper, rays, element_count, intensities, timestamps, ray_flags, batch_rays);
  integrate(*cpu.mapper, rays, element_count, intensities, timestamps, ray_flags, batch_rays);
  static_cast<ohm::GpuMap *>(gpu.mapper.get())->syncVoxels();
  if (regions_out)
  {
    *regions_out = gpu.map->regionCount();
  }
  if (timestamps && gpu.map->firstRayTime() != cpu.map->firstRayTime())
  {
    g_why = "firstRayTime differs";
    return -3;
  }
  return compareMaps(*gpu.map, *cpu.map, log_odds_tol, g_why);
}

// The map's callbacks into the cache (ohm/MapRegionCache.h) and the rest of the GpuMap surface.  0 = all good, else the
// number of the step that failed (binding_last_message says why).
int binding_cache_and_api(const double *rays, size_t element_count, size_t gpu_mem)
{
  g_why.clear();
  const auto *points = reinterpret_cast<const glm::dvec3 *>(rays);
  const unsigned flags = unsigned(ohm::MapFlag::kVoxelMean);
  const size_t half = (element_count / 4) * 2;

  // 1. a map that already holds data when the GpuMap is created: its chunks are uploaded, integration continues on them
  Side cpu = cpuSide(0, flags, 0.25);
  std::unique_ptr<ohm::OccupancyMap> map = makeMap(0, flags, 0.25);
  {
    ohm::RayMapperOccupancy first(map.get());
    first.integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  }
  cpu.mapper->integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  ohm::GpuMap gpu(map.get(), true, 2048u, gpu_mem);
  if (!gpu.gpuOk() || gpu.gpuCache() == nullptr || gpu.gpuCache() != ohm::gpumap::gpuCache(*map) || !gpu.borrowedMap() ||
      &gpu.map() != map.get())
  {
    g_why = "construction / gpuCache / map accessors";
    return 1;
  }
  gpu.integrateRays(points + half, element_count - half, nullptr, nullptr, ohm::kRfDefault);
  cpu.mapper->integrateRays(points + half, element_count - half, nullptr, nullptr, ohm::kRfDefault);
  gpu.syncVoxels();
  if (compareMaps(*map, *cpu.map, 0, g_why) != 0)
  {
    g_why = "upload of existing chunks + continue: " + g_why;
    return 1;
  }

  // 2. hit / miss values set through the mapper reach the device; syncVoxels(layers) and gpumap::sync
  gpu.setMissValue(-0.9f);
  gpu.setHitValue(1.1f);
  cpu.map->setMissValue(-0.9f);
  cpu.map->setHitValue(1.1f);
  if (gpu.missValue() != -0.9f || gpu.hitValue() != 1.1f)
  {
    g_why = "hit/miss accessors";
    return 2;
  }
  gpu.integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  cpu.mapper->integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  gpu.syncVoxels(std::vector<int>{ map->layout().occupancyLayer() });
  ohm::gpumap::sync(*map, unsigned(map->layout().meanLayer()));
  if (compareMaps(*map, *cpu.map, 0, g_why) != 0)
  {
    g_why = "set hit/miss + per-layer sync: " + g_why;
    return 2;
  }

  // 3. an arbitrary std::function filter (here: clip to 6 m): run on the host, clipped ends take a miss, order kept
  const ohm::RayFilterFunction clip = [](glm::dvec3 *s, glm::dvec3 *e, unsigned *f) { return ohm::clipRayFilter(s, e, f, 6.0); };
  gpu.setRayFilter(clip);
  cpu.map->setRayFilter(clip);
  gpu.integrateRays(points, element_count, nullptr, nullptr, ohm::kRfDefault);
  cpu.mapper->integrateRays(points, element_count, nullptr, nullptr, ohm::kRfDefault);
  gpu.syncVoxels();
  gpu.clearRayFilter();
  cpu.map->setRayFilter(ohm::OccupancyMap(1.0).rayFilter());
  if (compareMaps(*map, *cpu.map, 0, g_why) != 0)
  {
    g_why = "custom ray filter: " + g_why;
    return 3;
  }

  // 4. OccupancyMap::cullRegionsOutside -> MapRegionCache::remove for every culled region (OccupancyMap.cpp:1215-1217)
  ohm::GpuCache *cache = gpu.gpuCache();
  const size_t before = map->regionCount();
  map->cullRegionsOutside(glm::dvec3(-3.9), glm::dvec3(3.9));
  cpu.map->cullRegionsOutside(glm::dvec3(-3.9), glm::dvec3(3.9));
  const size_t after = map->regionCount();
  if (!(after < before) || ohmb200_region_count(cache->device()) != after)
  {
    g_why = "cull: host " + std::to_string(before) + " -> " + std::to_string(after) + ", device " +
            std::to_string(ohmb200_region_count(cache->device()));
    return 4;
  }
  gpu.integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  cpu.mapper->integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  gpu.syncVoxels();
  if (compareMaps(*map, *cpu.map, 0, g_why) != 0)
  {
    g_why = "integrate after cull: " + g_why;
    return 4;
  }

  // 5. OccupancyMap::clear -> MapRegionCache::clear (OccupancyMap.cpp:638); the pair keeps working afterwards
  map->clear();
  cpu.map->clear();
  if (ohmb200_region_count(cache->device()) != 0 || map->regionCount() != 0)
  {
    g_why = "clear did not reach the device";
    return 5;
  }
  gpu.integrateRays(points, element_count, nullptr, nullptr, ohm::kRfDefault);
  cpu.mapper->integrateRays(points, element_count, nullptr, nullptr, ohm::kRfDefault);
  gpu.syncVoxels();
  if (compareMaps(*map, *cpu.map, 0, g_why) != 0)
  {
    g_why = "integrate after clear: " + g_why;
    return 5;
  }

  // 6. findLayerCache / syncLayerTo (what ohm::copyMap asks of the cache, ohm/CopyUtil.cpp:103-107)
  const unsigned occ = unsigned(map->layout().occupancyLayer());
  if (cache->findLayerCache(occ) != cache || cache->findLayerCache(99) != nullptr)
  {
    g_why = "findLayerCache";
    return 6;
  }
  std::vector<const ohm::MapChunk *> chunks;
  map->enumerateRegions(chunks);
  gpu.integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);  // device now ahead of the host map
  cpu.mapper->integrateRays(points, half, nullptr, nullptr, ohm::kRfDefault);
  std::unique_ptr<ohm::OccupancyMap> copy = makeMap(0, flags, 0.25);
  for (const ohm::MapChunk *chunk : chunks)
  {
    ohm::MapChunk *dst = copy->region(chunk->region.coord, true);
    if (!cache->syncLayerTo(*dst, occ, *chunk, occ))
    {
      g_why = "syncLayerTo failed";
      return 6;
    }
    const ohm::MapChunk *expect = cpu.map->region(chunk->region.coord);
    ohm::VoxelBuffer<const ohm::VoxelBlock> a(dst->voxel_blocks[occ]), b(expect->voxel_blocks[occ]);
    if (memcmp(a.voxelMemory(), b.voxelMemory(), a.voxelMemorySize()) != 0)
    {
      g_why = "syncLayerTo: chunk differs from the CPU mapper's";
      return 6;
    }
  }
  return 0;
}

}  // extern "C"
