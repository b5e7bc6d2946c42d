// K1/K2 fast path: hierarchical conservative float32 frustum cull + exact float64 evaluation of the
// survivors, with the depth-map gathers software-pipelined one frame ahead.
//
// Why this is still bit-exact.  A point can only be visible (or splat into the z-buffer) if it lies in
// front of the camera and inside the colour image (info_handler.py:342-343,371).  With the standard
// intrinsics K = [fx 0 cx; 0 fy cy; 0 0 1] and Z > 0 that region is the intersection of four half-spaces
// of affine forms of the world point p (R0,R1,R2 = rows of w2c, "." includes the translation):
//     q0 =  i0 =  fx R0.p + cx R2.p          >= 0      (u >= 0)
//     q1 = -t0 = -fx R0.p + (W - cx) R2.p    >  0      (u <  W)
//     q2 =  i1 =  fy R1.p + cy R2.p          >= 0      (v >= 0)
//     q3 = -t1 = -fy R1.p + (H - cy) R2.p    >  0      (v <  H)
// (q0 >= 0 and q1 > 0 together imply Z > 0.)  The cull evaluates the forms in float32 and rejects a
// point only if some form is below  -bound_q, where bound_q = 16u * (un-cancelled magnitude of the form
// over the tile), u = 2^-24, is a rigorous bound on |float32 value - exact value| (coefficient rounding,
// coordinate rounding for float64 inputs, three FMA roundings: <= 5u).  A whole
// group of 32 consecutive points is rejected the same way through its bounding box (centre value +
// |a|.halfsize, evaluated with upward rounding).  Everything that survives — all in-frustum points plus
// a ~1e-2 pixel guard band — goes through the same exact float64 code as the reference-order kernel
// (msp_device.cuh), so the masks are identical to it; survivors are compacted per warp through a
// shared-memory queue so that the float64 work runs on dense warps whatever the point order.
#pragma once

#include "msp_device.cuh"

namespace msp {

constexpr int kCullThreads = 256;
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kCullPtsPerLane = 8;                             // 8 groups of 32 points per warp
constexpr int kCullWarpPoints = 32 * kCullPtsPerLane;          // 256
constexpr int kCullTilePoints = kCullWarps * kCullWarpPoints;  // 2048
constexpr int kCullFramesPerCta = 64;
constexpr int kCullPending = 4;  // survivor rounds per frame whose depth gather stays in flight across one frame

struct CullFrameConsts {  // per (CTA tile, frame); 96 bytes
  float form[4][4];       // {a0,a1,a2,d} of q0..q3 above, rounded to float32
  float bound[4];         // survive needs q >= -bound
  int std_frame;          // bottom row of w2c is exactly [0,0,0,1]
  int pad[3];
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

// monotone float <-> int maps for warp-wide min/max with REDUX
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

template <int DT, int MODE /*0 = visibility mask, 1 = z-buffer splat*/>
__global__ void __launch_bounds__(kCullThreads, 3) k_frame_cull(const CullParams p) {
  extern __shared__ __align__(128) unsigned char s_pts[];  // kCullTilePoints * PtTraits<DT>::kBytes (dynamic)
  __shared__ __align__(16) double s_w[kCullFramesPerCta * 16];
  __shared__ __align__(16) CullFrameConsts s_c[kCullFramesPerCta];
  __shared__ uint32_t s_words[kCullWarps][kCullPtsPerLane];
  __shared__ uint8_t s_q[2][kCullWarps][kCullWarpPoints];
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

  // ---- this lane's points (float32 registers), the tile maxima of |x|,|y|,|z| and the group boxes
  float px[kCullPtsPerLane], py[kCullPtsPerLane], pz[kCullPtsPerLane];
  const float qnan = __int_as_float(0x7fc00000);
  const float finf = __int_as_float(0x7f800000);
  float mx = 0.f, my = 0.f, mz = 0.f;
  float bcx = qnan, bcy = qnan, bcz = qnan, bhx = 0.f, bhy = 0.f, bhz = 0.f;  // box of group (lane & 7)
#pragma unroll
  for (int j = 0; j < kCullPtsPerLane; ++j) {
    const int idx = warp * kCullWarpPoints + j * 32 + lane;
    px[j] = py[j] = pz[j] = qnan;  // slots beyond the tile never survive the cull
    float lx = finf, ly = finf, lz = finf, hx = -finf, hy = -finf, hz = -finf;
    if (idx < cnt) {
      load_point_f32<DT>(s_pts, idx, px[j], py[j], pz[j]);
      if (isfinite(px[j]) && isfinite(py[j]) && isfinite(pz[j])) {
        lx = hx = px[j], ly = hy = py[j], lz = hz = pz[j];
        mx = fmaxf(mx, fabsf(px[j])), my = fmaxf(my, fabsf(py[j])), mz = fmaxf(mz, fabsf(pz[j]));
      }
    }
    const float gx0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lx))), gx1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hx)));
    const float gy0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(ly))), gy1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hy)));
    const float gz0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lz))), gz1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hz)));
    if ((lane & 7) == j) {
      // centre / half size with the half size rounded up so that [c-h, c+h] certainly contains the group
      // (for float64 inputs the float32 roundings of the coordinates are covered by the extra 2^-22 |c| term)
      const float k1 = 1.0f + 2.384185791015625e-07f, k2 = 4.76837158203125e-07f;  // 1+2^-22, 2^-21
      bcx = 0.5f * (gx0 + gx1), bcy = 0.5f * (gy0 + gy1), bcz = 0.5f * (gz0 + gz1);
      bhx = __fmaf_ru(0.5f * (gx1 - gx0), k1, k2 * fmaxf(fabsf(gx0), fabsf(gx1)));
      bhy = __fmaf_ru(0.5f * (gy1 - gy0), k1, k2 * fmaxf(fabsf(gy0), fabsf(gy1)));
      bhz = __fmaf_ru(0.5f * (gz1 - gz0), k1, k2 * fmaxf(fabsf(gz0), fabsf(gz1)));
      // an empty group has inf - inf = NaN in its centre: every test on it is false, the group is skipped
    }
  }
  {
    const uint32_t ux = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    const uint32_t uy = __reduce_max_sync(0xffffffffu, __float_as_uint(my));
    const uint32_t uz = __reduce_max_sync(0xffffffffu, __float_as_uint(mz));
    if (lane == 0) atomicMax(&s_max[0], ux), atomicMax(&s_max[1], uy), atomicMax(&s_max[2], uz);
  }
  __syncthreads();

  // ---- per-frame float32 forms and rigorous error bounds (one thread per frame)
  if (tid < nf) {
    const double* w = s_w + tid * 16;
    // a float64 coordinate rounded to float32 may exceed the true |x| by one ulp at most; inflate anyway
    const double kUp = 1.0 + 1.1920928955078125e-07;
    const double ax = (double)__uint_as_float(s_max[0]) * kUp, ay = (double)__uint_as_float(s_max[1]) * kUp,
                 az = (double)__uint_as_float(s_max[2]) * kUp;
    CullFrameConsts c;
    double S[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      S[r] = fabs(w[4 * r]) * ax + fabs(w[4 * r + 1]) * ay + fabs(w[4 * r + 2]) * az + fabs(w[4 * r + 3]);
    const double fx = p.K.k[0], cx = p.K.k[2], fy = p.K.k[5], cy = p.K.k[6], W = p.g.img_w, H = p.g.img_h;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c.form[0][k] = __double2float_rn(fx * w[k] + cx * w[8 + k]);
      c.form[1][k] = __double2float_rn(-fx * w[k] + (W - cx) * w[8 + k]);
      c.form[2][k] = __double2float_rn(fy * w[4 + k] + cy * w[8 + k]);
      c.form[3][k] = __double2float_rn(-fy * w[4 + k] + (H - cy) * w[8 + k]);
    }
    const double u16 = 16.0 * 5.9604644775390625e-08;  // 16 * 2^-24
    // un-cancelled magnitudes: they dominate sum_k |a_qk| m_k + |d_q| and also the reference's own float64 rounding
    c.bound[0] = __double2float_ru(u16 * (fabs(fx) * S[0] + fabs(cx) * S[2]));
    c.bound[1] = __double2float_ru(u16 * (fabs(fx) * S[0] + (fabs(W) + fabs(cx)) * S[2]));
    c.bound[2] = __double2float_ru(u16 * (fabs(fy) * S[1] + fabs(cy) * S[2]));
    c.bound[3] = __double2float_ru(u16 * (fabs(fy) * S[1] + (fabs(H) + fabs(cy)) * S[2]));
    // NaN / inf poses give NaN forms -> nothing survives, which is what the exact path returns for them too
    c.std_frame = (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
    if (!c.std_frame && isfinite(S[0] + S[1] + S[2])) {
      // perspective row present: the forms are not the image-plane tests, keep every point
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        c.form[q][0] = c.form[q][1] = c.form[q][2] = c.form[q][3] = 0.f;
        c.bound[q] = finf;
      }
    }
    c.pad[0] = c.pad[1] = c.pad[2] = 0;
    s_c[tid] = c;
  }
  __syncthreads();

  const int64_t word0 = (int64_t)blockIdx.x * (kCullTilePoints / 32) + warp * kCullPtsPerLane;
  const uint32_t lt = (1u << lane) - 1u;
  constexpr uint32_t kNoCand = 0xffffffffu;
  double zp[kCullPending];
  uint32_t rawp[kCullPending];
#pragma unroll
  for (int r = 0; r < kCullPending; ++r) zp[r] = 0.0, rawp[r] = kNoCand;
  uint32_t prev_total = 0;

  for (int fi = 0; fi <= nf; ++fi) {
    const int cur = fi & 1;
    uint32_t total = 0;
    // ---- stage A(fi): float32 cull, first per group box, then per point of the groups that are left
    if (fi < nf) {
      const CullFrameConsts& c = s_c[fi];
      const int q = lane >> 3;
      const float4 a = *reinterpret_cast<const float4*>(c.form[q]);
      const float v = fmaf(a.x, bcx, fmaf(a.y, bcy, fmaf(a.z, bcz, a.w)));
      const float r = __fmaf_ru(fabsf(a.x), bhx, __fmaf_ru(fabsf(a.y), bhy, __fmul_ru(fabsf(a.z), bhz)));
      const uint32_t gb = __ballot_sync(0xffffffffu, __fadd_ru(v, r) >= -c.bound[q]);
      const uint32_t gm = gb & (gb >> 8) & (gb >> 16) & (gb >> 24) & 0xffu;
      if (gm != 0u) {
        const float4 a0 = *reinterpret_cast<const float4*>(c.form[0]);
        const float4 a1 = *reinterpret_cast<const float4*>(c.form[1]);
        const float4 a2 = *reinterpret_cast<const float4*>(c.form[2]);
        const float4 a3 = *reinterpret_cast<const float4*>(c.form[3]);
        const float4 nb = *reinterpret_cast<const float4*>(c.bound);
#pragma unroll
        for (int j = 0; j < kCullPtsPerLane; ++j) {
          if (gm & (1u << j)) {  // warp-uniform
            const float q0 = fmaf(a0.x, px[j], fmaf(a0.y, py[j], fmaf(a0.z, pz[j], a0.w)));
            const float q1 = fmaf(a1.x, px[j], fmaf(a1.y, py[j], fmaf(a1.z, pz[j], a1.w)));
            const float q2 = fmaf(a2.x, px[j], fmaf(a2.y, py[j], fmaf(a2.z, pz[j], a2.w)));
            const float q3 = fmaf(a3.x, px[j], fmaf(a3.y, py[j], fmaf(a3.z, pz[j], a3.w)));
            const bool surv = (q0 >= -nb.x) & (q1 >= -nb.y) & (q2 >= -nb.z) & (q3 >= -nb.w);
            const uint32_t b = __ballot_sync(0xffffffffu, surv);
            if (surv) s_q[cur][warp][total + __popc(b & lt)] = (uint8_t)((j << 5) | lane);
            total += __popc(b);
          }
        }
      }
    }
    __syncwarp();

    // ---- consume(fi-1): the gathers issued one iteration ago have had stage A(fi) to land
    if (MODE == 0 && fi > 0) {
      const int64_t f = f0 + fi - 1;
#pragma unroll
      for (int r = 0; r < kCullPending; ++r) {
        if (rawp[r] != kNoCand) {
          if (depth_test(zp[r], (uint16_t)rawp[r], p.g)) {
            const uint32_t code = s_q[cur ^ 1][warp][r * 32 + lane];
            atomicOr(&s_words[warp][code >> 5], 1u << (code & 31));
          }
        }
      }
      __syncwarp();
      if (lane < kCullPtsPerLane) {
        const uint32_t wv = s_words[warp][lane];
        s_words[warp][lane] = 0u;
        if (word0 + lane < p.mask_stride) p.mask[f * p.mask_stride + word0 + lane] = wv;
      }
      __syncwarp();
    }

    // ---- stage B(fi): exact float64 projection of the compacted survivors; depth gathers are only issued here
    if (fi < nf) {
      const double* w = s_w + fi * 16;
      const bool fstd = s_c[fi].std_frame != 0;
      const int64_t f = f0 + fi;
#pragma unroll
      for (int r = 0; r < kCullPending; ++r) {
        rawp[r] = kNoCand;
        const uint32_t t = r * 32 + lane;
        if (t < total) {
          const uint32_t code = s_q[cur][warp][t];
          const int idx = warp * kCullWarpPoints + (int)(code >> 5) * 32 + (int)(code & 31);
          double x, y, z;
          PtTraits<DT>::load(s_pts, idx, x, y, z);
          const Proj pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
          if (in_image(pr, p.g) && pr.z > 0.0) {
            const int di = depth_index(pr, p.g);
            if (MODE == 0) {
              rawp[r] = ldg_u16_nc(p.depth + f * p.depth_stride + di);
              zp[r] = pr.z;
            } else {
              atomicMin(reinterpret_cast<unsigned int*>(p.zbuf + f * p.depth_stride) + di,
                        __float_as_uint(__double2float_rn(pr.z)));
            }
          }
        }
      }
      // more than kCullPending*32 survivors in this warp's 256 points: finish the rest synchronously
      for (uint32_t t0 = kCullPending * 32; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        if (t < total) {
          const uint32_t code = s_q[cur][warp][t];
          const int j = code >> 5, sl = code & 31;
          const int idx = warp * kCullWarpPoints + j * 32 + sl;
          double x, y, z;
          PtTraits<DT>::load(s_pts, idx, x, y, z);
          const Proj pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
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
    }
    prev_total = total;
    (void)prev_total;
  }
}

}  // namespace msp
