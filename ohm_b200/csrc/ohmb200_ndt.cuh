// ohmb200_ndt.cuh — NDT (normal distributions transform) voxel arithmetic, as the CPU mapper evaluates it.
//
// ohm/CovarianceVoxelCompute.h instantiates these with CovReal = double on the CPU (:33-34), storing the packed
// square-root covariance as six floats.  The device versions below keep the same operation order in fp64 with FMA
// contraction off, so the covariance update (sqrt, divide, multiply, add only) is bit-identical to the CPU's; the miss
// path additionally calls exp() and log(), whose last-ulp rounding differs between libm and CUDA — that is the stated
// tolerance on NDT log-odds (tests/test_gpu_ndt.py).
//   initialiseCovariance        CovarianceVoxelCompute.h:90-98
//   packedDot / unpackCovariance :101-163
//   solveTriangular             :183-204
//   calculateSampleLikelihoods  :227-267
//   calculateHitWithCovariance  :301-375
//   calculateMissNdt            :542-635
#pragma once

#include "ohmb200_device.cuh"

namespace ohmb200
{
OHMB200_HD __forceinline__ double dot3(const double a[3], const double b[3])
{
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];  // glm::dot
}

OHMB200_HD __forceinline__ double packedDot(const double A[9], int j, int k)
{
  const int indj = (j == 0) ? 0 : ((j == 1) ? 1 : 3);
  const int indk = (k == 0) ? 0 : ((k == 1) ? 1 : 3);
  const int m = (j <= k) ? j : k;
  double d = A[6 + k] * A[6 + j];
  for (int i = 0; i <= m; ++i)
  {
    d += A[indj + i] * A[indk + i];
  }
  return d;
}

OHMB200_HD __forceinline__ void solveTriangular(const float cov[6], const double y[3], double x[3])
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

// calculateSampleLikelihoods (CovarianceVoxelCompute.h:227-267): p(x_ML | N(mu, P)) and p(x_ML | z).
OHMB200_HD inline void ndtLikelihoods(const float cov[6], const double sensor[3], const double sample[3],
                                      const double mean[3], float sensor_noise, double &p_voxel, double &p_sample)
{
  double s2s[3], ray[3], m2s[3], a[3], bn[3], tmp[3], sol[3], x_ml[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    s2s[i] = sample[i] - sensor[i];
  }
  const double inv_len = 1.0 / sqrt(dot3(s2s, s2s));  // glm::normalize = v * inversesqrt(dot(v, v))
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    ray[i] = s2s[i] * inv_len;
    m2s[i] = sensor[i] - mean[i];
  }
  solveTriangular(cov, ray, a);
  solveTriangular(cov, m2s, bn);
  const double t = -dot3(a, bn) / dot3(a, a);
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    x_ml[i] = ray[i] * t + sensor[i];
    tmp[i] = x_ml[i] - mean[i];
  }
  solveTriangular(cov, tmp, sol);
  p_voxel = exp(-0.5 * dot3(sol, sol));
  const double noise_var = sensor_noise * sensor_noise;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    tmp[i] = x_ml[i] - sample[i];
  }
  p_sample = exp(-0.5 * dot3(tmp, tmp) / noise_var);
}

// Returns the log-odds adjustment of an NDT miss for a voxel with an established Gaussian (count >= threshold and
// observed); `valid` is false when the probability is NaN (the reference skips the update then).
OHMB200_HD inline float ndtMissAdjustment(const float cov[6], const double sensor[3], const double sample[3],
                                          const double mean[3], float adaptation_rate, float sensor_noise, bool &valid,
                                          bool &is_miss)
{
  double p_voxel, p_sample;
  ndtLikelihoods(cov, sensor, sample, mean, sensor_noise, p_voxel, p_sample);
  const double scaling = 0.5 * adaptation_rate;
  const double prod = p_voxel * (1.0 - p_sample);
  const double update = 0.5 - scaling * prod;
  is_miss = prod < scaling;
  valid = update == update;
  return valid ? (float)log(update / (1.0 - update)) : 0.0f;
}

// NDT-TM: calculateHitMissUpdateOnHit (CovarianceVoxelCompute.h:447-505), reinitialise_permeability_with_covariance = true
// (RayMapperNdt.cpp:325).  hm = {hit_count, miss_count}.
OHMB200_HD inline void ndtHitMissOnHit(const float cov[6], float value, uint2 &hm, const double sensor[3],
                                       const double sample[3], const double mean[3], uint32_t count, const MapParams &p)
{
  const bool needs_reset =
    value == INFINITY || (count == 0 || (value < p.reinit_threshold && count >= p.reinit_count));
  const uint32_t initial_hit = (!needs_reset) ? hm.x : 0;
  const uint32_t initial_miss = (!needs_reset) ? hm.y : 0;
  double p_voxel, p_sample;
  ndtLikelihoods(cov, sensor, sample, mean, p.sensor_noise, p_voxel, p_sample);
  const double prod = p_voxel * p_sample;
  const double eta = 0.5 * p.adaptation_rate;
  const bool inc_hit = needs_reset || count < p.sample_threshold || (count >= p.sample_threshold && prod >= eta);
  const bool inc_miss = !needs_reset && count >= p.sample_threshold && prod < eta && p_voxel >= eta;
  hm.x = initial_hit + (inc_hit ? 1u : 0u);
  hm.y = initial_miss + (inc_miss ? 1u : 0u);
}

// NDT-TM: calculateIntensityUpdateOnHit (CovarianceVoxelCompute.h:391-411).  im = {mean, covariance}.
OHMB200_HD inline void ndtIntensityOnHit(float2 &im, float value, float sample, uint32_t count, const MapParams &p)
{
  const bool needs_reset = count == 0 || (value < p.reinit_threshold && count >= p.reinit_count);
  const float delta = im.x - sample;
  const float n = (float)count;
  const float inv = 1.0f / (n + 1.0f);
  const float mean = (!needs_reset) ? inv * (n * im.x + sample) : sample;
  const float cv = (!needs_reset) ? inv * (n * im.y + inv * delta * delta) : p.initial_intensity_cov;
  im.x = mean;
  im.y = cv;
}

// calculateHitWithCovariance.  Returns true when the covariance was (re)initialised (the mean must restart).
OHMB200_HD inline bool ndtHit(float cov[6], float &value, const double sample[3], const double mean[3], uint32_t count,
                              float hit_value, float resolution, float reinit_threshold, uint32_t reinit_count)
{
  const float initial = value;
  const bool was_uncertain = initial == INFINITY;
  bool initialised = false;
  if (count == 0 || (initial < reinit_threshold && count >= reinit_count))
  {
    cov[0] = cov[2] = cov[5] = 0.1f * resolution;
    cov[1] = cov[3] = cov[4] = 0;
    initialised = true;
    count = 0;
  }
  value = (!was_uncertain) ? hit_value + initial : hit_value;

  double A[9];
  const double one_on = (double)1 / (count + (double)1);
  const double sc_1 = count ? sqrt(count * one_on) : (double)1;
  const double sc_2 = one_on * sqrt((double)count);
#pragma unroll
  for (int i = 0; i < 6; ++i)
  {
    A[i] = sc_1 * cov[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    A[6 + i] = sc_2 * ((!initialised) ? sample[i] - mean[i] : 0.0);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const int ind1 = (k * (k + 3)) >> 1;
    const int indk = ind1 - k;
    const double ak = sqrt(packedDot(A, k, k));
    cov[ind1] = (float)ak;
    if (ak > 0)
    {
      const double aki = (double)1 / ak;
#pragma unroll
      for (int j = k + 1; j < 3; ++j)
      {
        const int indj = (j * (j + 1)) >> 1;
        const int indkj = indj + k;
        double c = packedDot(A, j, k) * aki;
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

// ohm/VoxelOccupancyCompute.h:75-85 / :144-153 with null_update = false
OHMB200_HD __forceinline__ float adjustUp(float initial, float adjusted, const MapParams &p)
{
  const bool uninit = initial == INFINITY;
  adjusted = (uninit || (p.sat_min < initial && initial < p.sat_max)) ? adjusted : initial;
  return (adjusted != INFINITY) ? fminf(p.max_value, adjusted) : adjusted;
}

OHMB200_HD __forceinline__ float adjustDown(float initial, float adjusted, const MapParams &p)
{
  const bool uninit = initial == INFINITY;
  adjusted = (uninit || (p.sat_min < initial && initial < p.sat_max)) ? adjusted : initial;
  return (adjusted != INFINITY) ? fmaxf(p.min_value, adjusted) : adjusted;
}

// One full NDT miss on a voxel whose state is known (used when replaying ordered misses of a flagged voxel):
// calculateMissNdt + occupancyAdjustDown (RayMapperNdt.cpp:198-214).
OHMB200_HD inline float ndtMissOnce(float value, const float cov[6], const double sensor[3], const double sample[3],
                                    const double mean[3], uint32_t count, const MapParams &p, bool &is_miss)
{
  const float initial = value;
  float adjusted;
  is_miss = true;
  if (initial == INFINITY)
  {
    adjusted = p.miss_value;
  }
  else if (count < p.sample_threshold)
  {
    adjusted = initial + p.miss_value;
  }
  else
  {
    bool valid;
    const float adj = ndtMissAdjustment(cov, sensor, sample, mean, p.adaptation_rate, p.sensor_noise, valid, is_miss);
    adjusted = valid ? initial + adj : initial;
  }
  return adjustDown(initial, adjusted, p);
}

}  // namespace ohmb200
