// K1/K2 fast path: conservative float32 frustum cull + exact float64 evaluation of the survivors.
//
// Why this is still bit-exact.  A point can only be visible (or splat into the z-buffer) if it lies in
// front of the camera and inside the colour image (info_handler.py:342-343,371).  The cull evaluates
// the same affine forms in float32 and rejects a point only when a float32 value is outside the
// acceptance region by more than a rigorous bound on |float32 - exact| (derivation in DESIGN.md §K1):
//     X,Y,Z      : |Xf - X| <= 8u * S_X,  S_X = |e00| mx + |e01| my + |e02| mz + |e03|,  u = 2^-24
//     i0 = fx X + cx Z,  t0 = i0 - W Z  (u < 0  <=>  i0 < 0,  u >= W  <=>  t0 >= 0  when Z > 0)
//     bounds B = 16u * (fx S_X + cx S_Z) etc., per (tile, frame), mx/my/mz = tile maxima of |x|,|y|,|z|
// Everything that survives (all in-frustum points plus a ~1e-2 pixel guard band) goes through the same
// exact float64 code as the reference-order kernel (msp_device.cuh), so masks are identical to it.
// Survivors are compacted per warp through a shared-memory queue so the float64 work runs on dense
// warps regardless of the point order.
#pragma once

#include "msp_device.cuh"

namespace msp {

constexpr int kCullThreads = 256;
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kCullPtsPerLane = 8;                                   // 8 mask words per warp per frame
constexpr int kCullWarpPoints = 32 * kCullPtsPerLane;                // 256
constexpr int kCullTilePoints = kCullWarps * kCullWarpPoints;        // 2048
constexpr int kCullFramesPerCta = 64;

struct CullFrameConsts {  // per (CTA tile, frame); 80 bytes
  float e[12];            // w2c rows 0..2 rounded to float32
  float zmin;             // survive needs Zf > zmin            (= -8u S_Z)
  float b0, b0w;          // survive needs i0f >= -b0, t0f <= b0w
  float b1, b1h;          // same for v
  int std_frame;          // bottom row of w2c is exactly [0,0,0,1]
  float pad[2];
};

struct CullParams {
  const void* pts;
  int64_t n;
  const double* w2c;
  int32_t F;
  const uint16_t* depth;
  float* zbuf;
  int64_t depth_stride;
  uint32_t* mask;
  int64_t mask_stride;
  Intrinsics K;
  ImageGeom g;
  float fx, fy, cx, cy, fw, fh;  // float32 roundings of the intrinsics / image size (covered by the bounds)
};

template <int DT>
__device__ __forceinline__ void load_point_f32(const unsigned char* s_pts, int idx, float& x, float& y, float& z) {
  if (DT == 0) {
    const float* p = reinterpret_cast<const float*>(s_pts) + 3 * idx;
    x = p[0], y = p[1], z = p[2];
  } else if (DT == 1) {
    const float4 q = reinterpret_cast<const float4*>(s_pts)[idx];
    x = q.x, y = q.y, z = q.z;
  } else {
    const double* p = reinterpret_cast<const double*>(s_pts) + 3 * idx;
    x = __double2float_rn(p[0]), y = __double2float_rn(p[1]), z = __double2float_rn(p[2]);
  }
}

template <int DT, int MODE /*0 = visibility mask, 1 = z-buffer splat*/>
__global__ void __launch_bounds__(kCullThreads) k_frame_cull(const CullParams p) {
  extern __shared__ __align__(128) unsigned char s_pts[];  // kCullTilePoints * PtTraits<DT>::kBytes (dynamic)
  __shared__ __align__(16) double s_w[kCullFramesPerCta * 16];
  __shared__ __align__(16) CullFrameConsts s_c[kCullFramesPerCta];
  __shared__ uint32_t s_words[kCullWarps][kCullPtsPerLane];
  __shared__ uint8_t s_q[kCullWarps][kCullWarpPoints];
  __shared__ uint32_t s_max[3];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kCullTilePoints;
  const int cnt = (int)min((int64_t)kCullTilePoints, p.n - base);
  const int f0 = blockIdx.y * kCullFramesPerCta;
  const int nf = min(kCullFramesPerCta, p.F - f0);

  // ---- stage the point tile with one bulk async copy (TMA engine) + mbarrier
  const uint32_t bytes = (uint32_t)cnt * PtTraits<DT>::kBytes;
  const uint32_t bulk = bytes & ~15u;
  const unsigned char* gsrc = static_cast<const unsigned char*>(p.pts) + base * PtTraits<DT>::kBytes;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    s_max[0] = s_max[1] = s_max[2] = 0u;
  }
  if (tid < kCullWarps * kCullPtsPerLane) (&s_words[0][0])[tid] = 0u;
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s_bar, bulk);
    if (bulk) bulk_g2s(s_pts, gsrc, bulk, &s_bar);
  }
  if (tid < (int)(bytes - bulk)) s_pts[bulk + tid] = gsrc[bulk + tid];
  for (int i = tid; i < nf * 16; i += kCullThreads) s_w[i] = p.w2c[(int64_t)f0 * 16 + i];
  mbar_wait(&s_bar, 0);
  __syncthreads();

  // ---- this lane's points (float32 registers) and the tile maxima of |x|,|y|,|z|
  float px[kCullPtsPerLane], py[kCullPtsPerLane], pz[kCullPtsPerLane];
  const float qnan = __int_as_float(0x7fc00000);
  float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
  for (int j = 0; j < kCullPtsPerLane; ++j) {
    const int idx = warp * kCullWarpPoints + j * 32 + lane;
    px[j] = py[j] = pz[j] = qnan;  // slots beyond the tile never survive the cull
    if (idx < cnt) {
      load_point_f32<DT>(s_pts, idx, px[j], py[j], pz[j]);
      // the float32 rounding of a float64 coordinate can only round |x| up to the next float: take that as the max
      if (isfinite(px[j])) mx = fmaxf(mx, fabsf(px[j]));
      if (isfinite(py[j])) my = fmaxf(my, fabsf(py[j]));
      if (isfinite(pz[j])) mz = fmaxf(mz, fabsf(pz[j]));
    }
  }
  {
    const uint32_t ux = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    const uint32_t uy = __reduce_max_sync(0xffffffffu, __float_as_uint(my));
    const uint32_t uz = __reduce_max_sync(0xffffffffu, __float_as_uint(mz));
    if (lane == 0) atomicMax(&s_max[0], ux), atomicMax(&s_max[1], uy), atomicMax(&s_max[2], uz);
  }
  __syncthreads();

  // ---- per-frame float32 coefficients and rigorous error bounds (one thread per frame)
  if (tid < nf) {
    const double* w = s_w + tid * 16;
    // a float64 coordinate rounded to float32 may exceed the tile max by one ulp: inflate by (1 + 2^-23)
    const double kUp = 1.0 + 1.1920928955078125e-07;
    const double ax = (double)__uint_as_float(s_max[0]) * kUp, ay = (double)__uint_as_float(s_max[1]) * kUp,
                 az = (double)__uint_as_float(s_max[2]) * kUp;
    CullFrameConsts c;
    double S[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 4; ++k) c.e[4 * r + k] = __double2float_rn(w[4 * r + k]);
      S[r] = fabs(w[4 * r]) * ax + fabs(w[4 * r + 1]) * ay + fabs(w[4 * r + 2]) * az + fabs(w[4 * r + 3]);
    }
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double fx = fabs(p.K.k[0]), cx = fabs(p.K.k[2]), fy = fabs(p.K.k[5]), cy = fabs(p.K.k[6]);
    const double b0 = 16.0 * u * (fx * S[0] + cx * S[2]);
    const double b1 = 16.0 * u * (fy * S[1] + cy * S[2]);
    const double b0w = b0 + 16.0 * u * p.g.img_w * S[2];
    const double b1h = b1 + 16.0 * u * p.g.img_h * S[2];
    // round every bound upwards when it becomes float32; NaN / inf poses give NaN bounds -> nothing survives,
    // which is what the exact path returns for them too
    c.zmin = -__double2float_ru(8.0 * u * S[2]);
    c.b0 = __double2float_ru(b0), c.b0w = __double2float_ru(b0w);
    c.b1 = __double2float_ru(b1), c.b1h = __double2float_ru(b1h);
    c.std_frame = (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
    if (!c.std_frame && isfinite(S[0] + S[1] + S[2])) {
      // perspective row present: the affine forms above are not the image coordinates, keep every point
      const float inf = __int_as_float(0x7f800000);
      c.zmin = -inf, c.b0 = c.b0w = c.b1 = c.b1h = inf;
    }
    c.pad[0] = c.pad[1] = 0.f;
    s_c[tid] = c;
  }
  __syncthreads();

  const int64_t word0 = (int64_t)blockIdx.x * (kCullTilePoints / 32) + warp * kCullPtsPerLane;
  const uint32_t lt = (1u << lane) - 1u;

  for (int fi = 0; fi < nf; ++fi) {
    const CullFrameConsts& c = s_c[fi];
    const int64_t f = f0 + fi;
    // ---- stage A: float32 cull of this lane's 8 points
    uint32_t total = 0;
#pragma unroll
    for (int j = 0; j < kCullPtsPerLane; ++j) {
      const float X = fmaf(c.e[0], px[j], fmaf(c.e[1], py[j], fmaf(c.e[2], pz[j], c.e[3])));
      const float Y = fmaf(c.e[4], px[j], fmaf(c.e[5], py[j], fmaf(c.e[6], pz[j], c.e[7])));
      const float Z = fmaf(c.e[8], px[j], fmaf(c.e[9], py[j], fmaf(c.e[10], pz[j], c.e[11])));
      const float i0 = fmaf(p.cx, Z, p.fx * X);
      const float i1 = fmaf(p.cy, Z, p.fy * Y);
      const float t0 = fmaf(-p.fw, Z, i0);
      const float t1 = fmaf(-p.fh, Z, i1);
      const bool surv = (Z > c.zmin) & (i0 >= -c.b0) & (t0 <= c.b0w) & (i1 >= -c.b1) & (t1 <= c.b1h);
      const uint32_t b = __ballot_sync(0xffffffffu, surv);
      if (surv) s_q[warp][total + __popc(b & lt)] = (uint8_t)((j << 5) | lane);
      total += __popc(b);
    }
    __syncwarp();
    // ---- stage B: exact float64 evaluation of the compacted survivors
    const double* w = s_w + fi * 16;
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
      const uint32_t t = t0 + lane;
      if (t < total) {
        const uint32_t code = s_q[warp][t];
        const int j = code >> 5, sl = code & 31;
        const int idx = warp * kCullWarpPoints + j * 32 + sl;
        double x, y, z;
        PtTraits<DT>::load(s_pts, idx, x, y, z);
        const Proj pr = c.std_frame ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
        if (in_image(pr, p.g) && pr.z > 0.0) {
          const int di = depth_index(pr, p.g);
          if (MODE == 0) {
            if (depth_test(pr.z, ldg_u16_nc(p.depth + f * p.depth_stride + di), p.g)) atomicOr(&s_words[warp][j], 1u << sl);
          } else {
            atomicMin(reinterpret_cast<unsigned int*>(p.zbuf + f * p.depth_stride) + di,
                      __float_as_uint(__double2float_rn(pr.z)));
          }
        }
      }
    }
    if (MODE == 0) {
      __syncwarp();
      if (lane < kCullPtsPerLane) {
        const uint32_t wv = s_words[warp][lane];
        s_words[warp][lane] = 0u;
        if (word0 + lane < p.mask_stride) p.mask[f * p.mask_stride + word0 + lane] = wv;
      }
    }
    __syncwarp();
  }
}

}  // namespace msp
