// K1/K2 fast path (msp_project_visibility / msp_zbuffer_render with standard intrinsics).
//
// Structure per CTA: one tile of 2048 points (TMA bulk copy to SMEM) x one chunk of 64 frames.  Per frame and warp
// (256 points = 8 groups of 32 consecutive points, lane = point):
//   A  group classification on 32 lanes (4 half-space forms x 8 group boxes): groups certainly outside are skipped;
//   B  float32 evaluation of the remaining groups' points with rigorous error bounds: certainly outside -> done,
//      certainly inside with a certain depth pixel -> the depth gather is issued, anything uncertain -> exact queue;
//   C  float32 depth comparison with bounds (certain -> mask bit via ballot, uncertain -> exact queue);
//   D  the (rare) queued lanes are re-evaluated by the exact float64 reference-order code (msp_device.cuh).
// A float32 decision is only ever taken when the bound proves that the exact float64 reference arithmetic takes
// the same decision, so the output is bit-identical to the all-float64 kernel (k_frame_loop).  Proof sketch, u=2^-24:
//
//   Coordinates are re-centred on the tile centre c (x' = fl(x - c)); each affine form F(p) = a.p + d of
//   DESIGN.md §3 is evaluated as fl(a0 x' + fl(a1 y' + fl(a2 z' + d'))), d' = fl32(a.c + d) formed in float64.
//   (float64 inputs are re-centred in float64 and rounded once.)  Sources of error: coefficient rounding 1u,
//   re-centring / input rounding 1u (+1u slack), three FMA
//   roundings 3u on the a.x' part; 1u + 3u on d'.  Hence |F^ - F| <= 6u (A_F + |d'_F|), A_F = sum |a_k| ext_k,
//   ext_k = max_tile |x'_k|.  We use B_F = 8u (A_F + |d'_F|) + 2^-46 M_F (M_F = un-cancelled magnitude, covering the
//   float64 evaluation of d' and the reference's own float64 rounding).  Forms (already scaled to depth pixels,
//   sx = dw/W, sy = dh/H):  I0 = sx(fx X + cx Z),  Q1 = dw Z - I0,  I1 = sy(fy Y + cy Z),  Q3 = dh Z - I1,  Z.
//   In the image  <=>  I0 >= 0, Q1 > 0, I1 >= 0, Q3 > 0 (these imply Z > 0).  With B >= every form's bound
//   (derived forms Q1,Q3 included) and m = min(I0^,Q1^,I1^,Q3^):  m < -B  => outside in the reference;
//   m >= 2B => inside with margin >= B (enough to survive the reference's final division rounding);  else queue.
//   Pixel: t^ = I0^ * rcp(Z^) approximates the reference's u*sx within  dt = Ct1 * rcp(Z^) + Ct0,
//   Ct1 >= 1.01 (B_I0 + dw B_Z),  Ct0 >= 8u dw + 1e-9;  rint(t^) is the reference's pixel iff |t^ - rint(t^)| <= 1/2 - dt.
//   Depth: d^ = fl(float(raw) * 0.001f) is within 2.1u d of raw*0.001;  Z^ within B_Z of the reference's z;
//   visible iff z < d is certain when Z^ < d^ - (B_Z + 6u d^), certainly false when Z^ >= d^ + (B_Z + 6u d^).
#pragma once

#include "msp_device.cuh"

namespace msp {

constexpr int kCullThreads = 256;
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kCullPtsPerLane = 8;                             // 8 groups of 32 points per warp
constexpr int kCullWarpPoints = 32 * kCullPtsPerLane;          // 256
constexpr int kCullTilePoints = kCullWarps * kCullWarpPoints;  // 2048
#ifndef MSP_CULL_FRAMES
#define MSP_CULL_FRAMES 48
#endif
constexpr int kCullFramesPerCta = MSP_CULL_FRAMES;
constexpr int kCullMaxObjects = 63;
#ifndef MSP_CULL_MINBLOCKS
#define MSP_CULL_MINBLOCKS 4
#endif
#ifndef MSP_CULL_PREFETCH
#define MSP_CULL_PREFETCH 0  // 1: each CTA prefetches its share of the chunk's depth maps into L2 while it sets up (measured: no gain, profiles/tuning_r1.md)
#endif
#ifndef MSP_CULL_BOX_SMEM
#define MSP_CULL_BOX_SMEM 1    // 1: the warp's 8 group boxes live in SMEM (2 LDS.128 per frame) instead of 6 registers, two of which spilled
#endif
#ifndef MSP_CULL_DEFER_EXACT
#define MSP_CULL_DEFER_EXACT 1     // 1 (visibility mode): lanes queued for the exact float64 path are collected per warp ACROSS frames and
                                   // evaluated 32 at a time, OR-ing the few visible bits into the already stored mask words, instead of
                                   // running the ~200-instruction exact path for one or two lanes in ~15 % of the warp-frames
                                   // (end of round 1: -10 % on a 200 k x 512 scene with bit-identical masks and counts,
                                   // profiles/variant_check_r1.txt; round 2: whole -m gpu suite green on it, shipped on)
#endif
#ifndef MSP_CULL_STRIDED_GROUPS
#define MSP_CULL_STRIDED_GROUPS 1  // 1: warp w owns the tile's groups w, w+8, w+16, ... instead of 8 consecutive ones, so that the eight
                                   // warps of a CTA sample the tile's extent alike and finish their frames together (end of round 1:
                                   // -2.6 %, bit-identical output, profiles/variant_check_r1.txt; round 2: whole -m gpu suite green, shipped on)
#endif
#if MSP_CULL_STRIDED_GROUPS
#define MSP_CULL_GROUP(warp, j) ((j) * kCullWarps + (warp))
#else
#define MSP_CULL_GROUP(warp, j) ((warp) * kCullPtsPerLane + (j))
#endif
#ifndef MSP_CULL_BOX_ADDR
#define MSP_CULL_BOX_ADDR 1    // 1 (with BOX_SMEM): the lane's box address is kept in a register (opaque) instead of being re-derived every frame
#endif
#ifndef MSP_CULL_PIPELINE
#define MSP_CULL_PIPELINE 0    // 1: the depth gathers of a pair of groups are consumed one loop iteration later (measured slower: profiles/tuning_r1.md)
#endif
#ifndef MSP_CULL_MAGIC_RINT
#define MSP_CULL_MAGIC_RINT 1  // 1: rint(t) as (t + 1.5*2^23) - 1.5*2^23 on the FMA pipe instead of FRND on the XU pipe (exact for |t| < 2^22)
#endif
#ifndef MSP_CULL_MAGIC_I2F
#define MSP_CULL_MAGIC_I2F 1   // 1: float(raw u16) as as_float(0x4B000000 | raw) - 2^23 instead of I2F on the XU pipe (exact)
#endif
#ifndef MSP_ZBUF_WARP_MIN
#define MSP_ZBUF_WARP_MIN 0  // K2, 1: warp-level nearest-z agreement per depth pixel (MATCH.ANY + REDUX.MIN) before the global atomicMin (measured slower: profiles/tuning_r1.md)
#endif
#ifndef MSP_CULL_W_SMEM
#define MSP_CULL_W_SMEM 0  // 1: keep the chunk's float64 poses in SMEM for the exact path, 0: read them from global
#endif  // fused object counting keeps a (64 frames x 64 ids) histogram in SMEM

struct CullFrameConsts {  // per (CTA tile, frame); 128 bytes, all relative to the tile centre
  float box[4][4];        // I0, Q1, I1, Q3 as {a0,a1,a2,d'} for the group-box test
  float fz[4];            // Z form {a0,a1,a2,d'}
  float B;                // common bound of the four in-image forms
  float Bz;               // bound of the Z form
  float ct1, ct0;         // pixel-index guard band  dt = ct1 * rcp(Z) + ct0
  int std_frame;          // bottom row of w2c is exactly [0,0,0,1]
  int tile_cand;          // 0: the whole 2048-point tile is certainly outside this frame's image
  unsigned long long dm_biased;  // this frame's depth map as a byte address minus 2 * 0x4B000000 (see eval32's gather)
  int pad[4];
};
static_assert(sizeof(CullFrameConsts) == 128, "CullFrameConsts layout");

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
  float fdw, fdh, fdw1, fdh1;  // depth_w, depth_h, depth_w-1, depth_h-1 as float32 (exact), kept in the constant bank
  float fscale;               // float32 rounding of depth_scale (its error is inside the 6u d guard)
  const uint16_t* inst;       // optional (n): 1-indexed instance ids for fused per-(frame, object) counts
  uint32_t* counts;           // optional (F, n_objects), zeroed by the caller; accumulated with atomics
  int32_t n_objects;          // <= kCullMaxObjects when counts != NULL
  unsigned long long* stats;  // optional: [0] lanes evaluated in float32, [1] lanes sent to the exact path
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
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // max relative error 2^-23 (PTX ISA), covered by ct0/ct1
  return r;
}

template <int DT, int MODE /*0 = visibility mask, 1 = z-buffer splat*/, bool COUNT = false>
__global__ void __launch_bounds__(kCullThreads, MSP_CULL_MINBLOCKS) k_frame_cull(const CullParams p) {
  // per (frame, id) visible counts of this tile: <= 2048 each, two 16-bit counters per word
  constexpr int kCntWords = (kCullMaxObjects + 1) / 2;
  __shared__ uint32_t s_cnt[COUNT ? kCullFramesPerCta * kCntWords : 1];
  extern __shared__ __align__(128) unsigned char s_pts[];  // kCullTilePoints * PtTraits<DT>::kBytes (dynamic)
#if MSP_CULL_W_SMEM
  __shared__ __align__(16) double s_w[kCullFramesPerCta * 16];
#endif
  __shared__ __align__(16) CullFrameConsts s_c[kCullFramesPerCta];
#if MSP_CULL_BOX_SMEM
  __shared__ __align__(16) float s_box[kCullWarps][kCullPtsPerLane][8];  // per group: centre xyz, -, half size xyz, - (re-centred)
#endif
  __shared__ uint32_t s_words[kCullWarps][kCullPtsPerLane];
  __shared__ uint8_t s_q[kCullWarps][kCullWarpPoints];
#if MSP_CULL_DEFER_EXACT
  constexpr uint32_t kDeferCap = 64;                        // fill < 32 between frames, at most 32 appended per frame
  __shared__ uint16_t s_dq[kCullWarps][kDeferCap];          // (frame in chunk << 8) | (group << 5) | lane
#endif
  __shared__ uint8_t s_flist[kCullFramesPerCta];  // frames of this chunk in which the tile may be visible
  __shared__ int s_nlist;
  __shared__ int s_red[9];  // ordered-int min xyz, max xyz, then uint max |x'|,|y'|,|z'|
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
    s_red[0] = s_red[1] = s_red[2] = 0x7fffffff;
    s_red[3] = s_red[4] = s_red[5] = (int)0x80000000;
    s_red[6] = s_red[7] = s_red[8] = 0;
  }
  if (tid < kCullWarps * kCullPtsPerLane) (&s_words[0][0])[tid] = 0u;
  if (COUNT)
    for (int i = tid; i < kCullFramesPerCta * kCntWords; i += kCullThreads) s_cnt[i] = 0u;
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s_bar, bulk);
    if (bulk) bulk_g2s(s_pts, gsrc, bulk, &s_bar);
  }
  if (tid < (int)(bytes - bulk)) s_pts[bulk + tid] = gsrc[bulk + tid];
#if MSP_CULL_PREFETCH
  if (MODE == 0 && warp == 1) {
    // The tiles of one frame chunk are co-resident (blockIdx.x is the fast grid dimension) and all of them gather from
    // the same nf depth maps a few microseconds from now: pull this CTA's 1/gridDim.x share of those maps into L2 with
    // bulk prefetches so that the gathers find them there instead of paying the HBM latency.
    const uint64_t total = (uint64_t)nf * (uint64_t)p.depth_stride * sizeof(uint16_t);
    const uint64_t slice = ((total + gridDim.x - 1) / gridDim.x + 127u) & ~(uint64_t)127u;
    const uint64_t lo = (uint64_t)blockIdx.x * slice;
    const uint64_t hi = (lo + slice < total) ? lo + slice : total;
    const char* dbase = reinterpret_cast<const char*>(p.depth + (int64_t)f0 * p.depth_stride);
    constexpr uint32_t kPiece = 16384;
    for (uint64_t o = lo + (uint64_t)lane * kPiece; o < hi; o += 32ull * kPiece) {
      const uint32_t sz = (uint32_t)((hi - o < kPiece) ? (hi - o) : kPiece) & ~15u;
      if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(dbase + o), "r"(sz) : "memory");
    }
  }
#endif
#if MSP_CULL_W_SMEM
  for (int i = tid; i < nf * 16; i += kCullThreads) s_w[i] = p.w2c[(int64_t)f0 * 16 + i];
#else
  const double* s_w = p.w2c + (int64_t)f0 * 16;  // read through L2 by the per-frame prep threads and the rare exact path
#endif
  mbar_wait(&s_bar, 0);
  __syncthreads();

  // ---- fused object counting: instance id of each of this lane's points; groups whose 32 ids agree are counted
  //      with one popcount (neighbouring mesh vertices mostly belong to one object, compute_object_visibility.py:110-136)
  uint32_t ids_lo = 0, ids_hi = 0, ids_lo2 = 0, ids_hi2 = 0;  // 8 x u16 packed
  uint32_t uni_mask = 0;                                       // bit j: group j has one id (kept in gid of lane j)
  uint32_t gid = 0;
  if (COUNT) {
#pragma unroll
    for (int j = 0; j < kCullPtsPerLane; ++j) {
      const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
      uint32_t id = (idx < cnt) ? (uint32_t)__ldg(p.inst + base + idx) : 0u;
      if (id > (uint32_t)p.n_objects) id = 0u;
      const uint32_t first = __shfl_sync(0xffffffffu, id, 0);
      // slots beyond the tile never become visible, so they may be ignored when testing uniformity
      const bool same = __all_sync(0xffffffffu, id == first || idx >= cnt);
      if (same) uni_mask |= 1u << j;
      if (lane == j) gid = first;
      const uint32_t sh = (uint32_t)(j & 1) * 16u;
      if (j < 2) ids_lo |= id << sh; else if (j < 4) ids_hi |= id << sh; else if (j < 6) ids_lo2 |= id << sh; else ids_hi2 |= id << sh;
    }
  }

  // ---- this lane's points, tile box -> centre, re-centred coordinates, group boxes, extents
  float px[kCullPtsPerLane], py[kCullPtsPerLane], pz[kCullPtsPerLane];
  const float qnan = __int_as_float(0x7fc00000);
  const float finf = __int_as_float(0x7f800000);
  {
    float lx = finf, ly = finf, lz = finf, hx = -finf, hy = -finf, hz = -finf;
#pragma unroll
    for (int j = 0; j < kCullPtsPerLane; ++j) {
      const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
      px[j] = py[j] = pz[j] = qnan;  // slots beyond the tile and non-finite points: every comparison on them is false
      if (idx < cnt) {
        float x, y, z;
        load_point_f32<DT>(s_pts, idx, x, y, z);
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
          px[j] = x, py[j] = y, pz[j] = z;
          lx = fminf(lx, x), ly = fminf(ly, y), lz = fminf(lz, z);
          hx = fmaxf(hx, x), hy = fmaxf(hy, y), hz = fmaxf(hz, z);
        }
      }
    }
    const int ox0 = __reduce_min_sync(0xffffffffu, f2ord(lx)), ox1 = __reduce_max_sync(0xffffffffu, f2ord(hx));
    const int oy0 = __reduce_min_sync(0xffffffffu, f2ord(ly)), oy1 = __reduce_max_sync(0xffffffffu, f2ord(hy));
    const int oz0 = __reduce_min_sync(0xffffffffu, f2ord(lz)), oz1 = __reduce_max_sync(0xffffffffu, f2ord(hz));
    if (lane == 0) {
      atomicMin(&s_red[0], ox0), atomicMin(&s_red[1], oy0), atomicMin(&s_red[2], oz0);
      atomicMax(&s_red[3], ox1), atomicMax(&s_red[4], oy1), atomicMax(&s_red[5], oz1);
    }
  }
  __syncthreads();
  // tile centre (any finite value works for correctness; an empty tile gets 0)
  float tcx = 0.5f * (ord2f(s_red[0]) + ord2f(s_red[3])), tcy = 0.5f * (ord2f(s_red[1]) + ord2f(s_red[4])),
        tcz = 0.5f * (ord2f(s_red[2]) + ord2f(s_red[5]));
  if (!isfinite(tcx) || !isfinite(tcy) || !isfinite(tcz)) tcx = tcy = tcz = 0.f;
  float bcx = qnan, bcy = qnan, bcz = qnan, bhx = 0.f, bhy = 0.f, bhz = 0.f;  // box of group (lane & 7), re-centred
  // From here on the tile buffer holds the RE-CENTRED float32 coordinates (x',y',z' packed, NaN for empty slots);
  // the original coordinates are only needed by the rare exact path, which reads them from global memory (L2).
  float* s_rel = reinterpret_cast<float*>(s_pts);
  {
    float ex = 0.f, ey = 0.f, ez = 0.f;
#pragma unroll
    for (int j = 0; j < kCullPtsPerLane; ++j) {
      if (DT == 2) {
        // float64 inputs: subtract the centre in float64 and round ONCE, so that the error of x' is u |x'| and not
        // u |x| (rounding the coordinate itself to float32 first loses |x| * 2^-24: metres at kilometre offsets)
        if (px[j] == px[j]) {
          const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
          double x64, y64, z64;
          PtTraits<DT>::load(s_pts, idx, x64, y64, z64);
          px[j] = __double2float_rn(x64 - (double)tcx);
          py[j] = __double2float_rn(y64 - (double)tcy);
          pz[j] = __double2float_rn(z64 - (double)tcz);
        }
      } else {
        px[j] -= tcx, py[j] -= tcy, pz[j] -= tcz;  // NaN stays NaN
      }
      const bool ok = px[j] == px[j];
      ex = fmaxf(ex, fabsf(px[j])), ey = fmaxf(ey, fabsf(py[j])), ez = fmaxf(ez, fabsf(pz[j]));  // fmaxf ignores NaN
      const float lx = ok ? px[j] : finf, ly = ok ? py[j] : finf, lz = ok ? pz[j] : finf;
      const float hx = ok ? px[j] : -finf, hy = ok ? py[j] : -finf, hz = ok ? pz[j] : -finf;
      const float gx0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lx))), gx1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hx)));
      const float gy0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(ly))), gy1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hy)));
      const float gz0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lz))), gz1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hz)));
      if ((lane & 7) == j) {
        // centre / half size with the half size rounded up so that [c-h, c+h] certainly contains the group's x'
        const float k1 = 1.0f + 2.384185791015625e-07f, k2 = 4.76837158203125e-07f;  // 1+2^-22, 2^-21
        bcx = 0.5f * (gx0 + gx1), bcy = 0.5f * (gy0 + gy1), bcz = 0.5f * (gz0 + gz1);
        bhx = __fmaf_ru(0.5f * (gx1 - gx0), k1, k2 * fmaxf(fabsf(gx0), fabsf(gx1)));
        bhy = __fmaf_ru(0.5f * (gy1 - gy0), k1, k2 * fmaxf(fabsf(gy0), fabsf(gy1)));
        bhz = __fmaf_ru(0.5f * (gz1 - gz0), k1, k2 * fmaxf(fabsf(gz0), fabsf(gz1)));
        // an empty group has inf - inf = NaN in its centre: every test on it is false, the group is skipped
      }
    }
#if MSP_CULL_BOX_SMEM
    if (lane < kCullPtsPerLane) {
      *reinterpret_cast<float4*>(&s_box[warp][lane][0]) = make_float4(bcx, bcy, bcz, 0.f);
      *reinterpret_cast<float4*>(&s_box[warp][lane][4]) = make_float4(bhx, bhy, bhz, 0.f);
    }
#endif
    const uint32_t ux = __reduce_max_sync(0xffffffffu, __float_as_uint(ex));
    const uint32_t uy = __reduce_max_sync(0xffffffffu, __float_as_uint(ey));
    const uint32_t uz = __reduce_max_sync(0xffffffffu, __float_as_uint(ez));
    if (lane == 0) atomicMax(&s_red[6], (int)ux), atomicMax(&s_red[7], (int)uy), atomicMax(&s_red[8], (int)uz);
  }
  __syncthreads();  // every thread has read its original coordinates: the buffer may now be overwritten
#pragma unroll
  for (int j = 0; j < kCullPtsPerLane; ++j) {
    const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
    s_rel[3 * idx + 0] = px[j], s_rel[3 * idx + 1] = py[j], s_rel[3 * idx + 2] = pz[j];
  }

  // ---- per-frame float32 forms and rigorous error bounds (one thread per frame, float64 arithmetic)
  if (tid < nf) {
    const double* w = s_w + tid * 16;
    const double kUp = 1.0 + 2.4e-07;  // extents: cover the rounding of x' itself and of float64 inputs
    const double ext[3] = {(double)__int_as_float(s_red[6]) * kUp, (double)__int_as_float(s_red[7]) * kUp,
                           (double)__int_as_float(s_red[8]) * kUp};
    const double tc[3] = {(double)tcx, (double)tcy, (double)tcz};
    const double fx = p.K.k[0], cx = p.K.k[2], fy = p.K.k[5], cy = p.K.k[6];
    const double sx = p.g.sx, sy = p.g.sy, dw = (double)p.g.depth_w, dh = (double)p.g.depth_h;
    // float64 coefficient rows of I0, Q1, I1, Q3, Z on original coordinates (k = 0..2 linear, 3 = constant)
    double A[5][4], M[5];
    double mag[3];  // un-cancelled |row| . (|c| + ext) + |t| of the three w2c rows
#pragma unroll
    for (int r = 0; r < 3; ++r)
      mag[r] = fabs(w[4 * r]) * (fabs(tc[0]) + ext[0]) + fabs(w[4 * r + 1]) * (fabs(tc[1]) + ext[1]) +
               fabs(w[4 * r + 2]) * (fabs(tc[2]) + ext[2]) + fabs(w[4 * r + 3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[0][k] = sx * (fx * w[k] + cx * w[8 + k]);
      A[2][k] = sy * (fy * w[4 + k] + cy * w[8 + k]);
      A[4][k] = w[8 + k];
      A[1][k] = dw * w[8 + k] - A[0][k];
      A[3][k] = dh * w[8 + k] - A[2][k];
    }
    M[0] = fabs(sx) * (fabs(fx) * mag[0] + fabs(cx) * mag[2]);
    M[2] = fabs(sy) * (fabs(fy) * mag[1] + fabs(cy) * mag[2]);
    M[4] = mag[2];
    M[1] = dw * mag[2] + M[0];
    M[3] = dh * mag[2] + M[2];
    CullFrameConsts c;
    const double u = 5.9604644775390625e-08, tiny = 1.4210854715202004e-14;  // 2^-24, 2^-46
    double Bf[5], top[4];  // top[q]: upper bound of form q over the whole tile (centre value + |a|.extent)
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double dprime = A[q][0] * tc[0] + A[q][1] * tc[1] + A[q][2] * tc[2] + A[q][3];
      const double Aext = fabs(A[q][0]) * ext[0] + fabs(A[q][1]) * ext[1] + fabs(A[q][2]) * ext[2];
      if (q < 4) top[q] = dprime + Aext;
      Bf[q] = 8.0 * u * (Aext + fabs(dprime)) + tiny * M[q];
      float* dst = (q < 4) ? c.box[q] : c.fz;
      dst[0] = __double2float_rn(A[q][0]), dst[1] = __double2float_rn(A[q][1]), dst[2] = __double2float_rn(A[q][2]);
      dst[3] = __double2float_rn(dprime);
    }
    // Q1 and Q3 are *derived* in the point loop as fl(dw Z^ - I0^): their error is B_I0 + dw B_Z + one rounding
    // (the final FMA rounding u |Q^| <= u (A_ext + |d'|)(1+..) <= 0.13 Bf of the directly evaluated form)
    const double bq1 = 1.01 * (Bf[0] + dw * Bf[4]) + 0.13 * Bf[1];
    const double bq3 = 1.01 * (Bf[2] + dh * Bf[4]) + 0.13 * Bf[3];
    const double B = fmax(fmax(Bf[0], Bf[2]), fmax(fmax(Bf[1], Bf[3]), fmax(bq1, bq3)));
    c.B = __double2float_ru(B);
    c.Bz = __double2float_ru(Bf[4]);
    c.ct1 = __double2float_ru(1.01 * fmax(Bf[0] + dw * Bf[4], Bf[2] + dh * Bf[4]));
    c.ct0 = __double2float_ru(8.0 * u * fmax(dw, dh) + 1e-9);
    c.std_frame = (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
    // every point of the tile satisfies form_q <= top[q] (exact arithmetic); if that is below -B for some q the whole
    // tile is outside.  NaN poses: comparisons false -> tile_cand = 0, and indeed nothing is visible.
    c.tile_cand = (top[0] >= -B) && (top[1] >= -B) && (top[2] >= -B) && (top[3] >= -B);
    if (!c.std_frame && isfinite(B)) {
      c.tile_cand = 1;
      // perspective row present: the forms are not the image-plane tests — send every point to the exact path
#pragma unroll
      for (int q = 0; q < 4; ++q) c.box[q][0] = c.box[q][1] = c.box[q][2] = c.box[q][3] = 0.f;
      c.B = finf;
    }
    c.dm_biased = (unsigned long long)(p.depth + (int64_t)(f0 + tid) * p.depth_stride) - 2ull * 0x4B000000ull;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.pad[i] = 0;
    s_c[tid] = c;
  }
  __syncthreads();

#if MSP_CULL_STRIDED_GROUPS
  // lane j (< 8) owns the mask word of the warp's group j: word0 + lane is that word's index
  const int64_t word0 = (int64_t)blockIdx.x * (kCullTilePoints / 32) + MSP_CULL_GROUP(warp, lane & (kCullPtsPerLane - 1)) - lane;
#else
  const int64_t word0 = (int64_t)blockIdx.x * (kCullTilePoints / 32) + warp * kCullPtsPerLane;
#endif
  const uint32_t lt = (1u << lane) - 1u;
  const bool word_ok = (lane < kCullPtsPerLane) && (word0 + lane < p.mask_stride);
  uint32_t* mask_ptr = p.mask + (int64_t)f0 * p.mask_stride + word0 + lane;
  // this lane's point of group j is at shared address s_lane + 384 j; the address is made opaque so that it is kept in a
  // register instead of being re-derived from the thread id at every use (8 instructions per point at 64 registers)
  constexpr uint32_t kGroupStrideBytes = (uint32_t)(MSP_CULL_GROUP(0, 1) - MSP_CULL_GROUP(0, 0)) * 32u * 12u;  // 384 / 3072
  uint32_t s_lane = (uint32_t)__cvta_generic_to_shared(s_rel + 3 * (MSP_CULL_GROUP(warp, 0) * 32 + lane));
  asm volatile("" : "+r"(s_lane));
#if MSP_CULL_BOX_SMEM && MSP_CULL_BOX_ADDR
  uint32_t s_box_lane = (uint32_t)__cvta_generic_to_shared(&s_box[warp][lane & 7][0]);
  asm volatile("" : "+r"(s_box_lane));
#endif
  unsigned int st_f32 = 0, st_exact = 0;

  // float32 evaluation of one group's point on this lane (branch-free; see the header comment for the bounds)
  struct Lane32 {
    float zf;
    uint32_t raw;
    bool go, unc;
  };
  const double* wz = nullptr;      // MODE 1: row 2 of the current frame's w2c (float64) and its z-buffer row
  unsigned int* zrow = nullptr;
  auto eval32 = [&](int j, const float4& a0, const float4& a2, const float4& az, float B, float B2, float ct1,
                    float half_ct0, unsigned long long dmb) -> Lane32 {
    float x, y, z;
    {
      const uint32_t q = s_lane + kGroupStrideBytes * (uint32_t)j;
      asm("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(q));
      asm("ld.shared.f32 %0, [%1+4];" : "=f"(y) : "r"(q));
      asm("ld.shared.f32 %0, [%1+8];" : "=f"(z) : "r"(q));
    }
    const float I0 = fmaf(a0.x, x, fmaf(a0.y, y, fmaf(a0.z, z, a0.w)));
    const float I1 = fmaf(a2.x, x, fmaf(a2.y, y, fmaf(a2.z, z, a2.w)));
    const float Z = fmaf(az.x, x, fmaf(az.y, y, fmaf(az.z, z, az.w)));
    const float Q1 = fmaf(p.fdw, Z, -I0), Q3 = fmaf(p.fdh, Z, -I1);
    const float m = fminf(fminf(I0, Q1), fminf(I1, Q3));  // NaN only for empty / non-finite slots: never visible
    const bool cin = m >= B2;                               // certainly inside the image, in front of the camera
    const float rz = rcp_approx(Z);
    const float tx = I0 * rz, ty = I1 * rz;
#if MSP_CULL_MAGIC_RINT
    const float ix = (tx + 12582912.0f) - 12582912.0f, iy = (ty + 12582912.0f) - 12582912.0f;  // = rintf for |t| < 2^22; lanes beyond are never `go`
#else
    const float ix = rintf(tx), iy = rintf(ty);
#endif
    const float e = fmaxf(fabsf(tx - ix), fabsf(ty - iy));
    const bool pc = e <= fmaf(-ct1, rz, half_ct0);          // the reference rounds to the same depth pixel
    const float di = fmaf(fminf(iy, p.fdh1), p.fdw, fminf(ix, p.fdw1));
    Lane32 r;
    r.go = cin && pc;
    r.unc = (m >= -B) && !r.go;  // not provably outside and not decided here
    r.zf = Z;
    r.raw = 0u;
    if (MODE == 1) {
      // z-buffer splat: the pixel is certain, the stored value must be float32(z_ref) — recompute z alone in float64
      // from the original coordinates (IH:60,63 operation order) and splat; nothing is left to compare later
#if MSP_ZBUF_WARP_MIN
      // lanes of the warp that hit the same depth pixel first agree on their nearest z (positive floats order like their
      // bit patterns): MATCH.ANY groups them by pixel, REDUX.MIN reduces inside each group, and only the lanes holding the
      // group's minimum touch memory - one atomic per (warp, pixel) instead of one per point
      const uint32_t act = __ballot_sync(0xffffffffu, r.go);
      if (r.go) {
        const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
        double x64, y64, z64;
        PtTraits<DT>::load(p.pts, base + idx, x64, y64, z64);
        const uint32_t zb = __float_as_uint(__double2float_rn(row_dot_h(wz, x64, y64, z64)));
        const uint32_t peers = __match_any_sync(act, (int)di);
        uint32_t zmin = zb;
        if (peers & (peers - 1u)) zmin = __reduce_min_sync(peers, zb);   // uniform within a group of peers
        if (zb == zmin) atomicMin(zrow + (int)di, zb);
      }
#else
      if (r.go) {
        const int idx = MSP_CULL_GROUP(warp, j) * 32 + lane;
        double x64, y64, z64;
        PtTraits<DT>::load(p.pts, base + idx, x64, y64, z64);
        const double zref = row_dot_h(wz, x64, y64, z64);
        atomicMin(zrow + (int)di, __float_as_uint(__double2float_rn(zref)));
      }
#endif
      r.go = false;
      return r;
    }
    // di is an exact integer below 2^23 on every lane that gathers: adding 2^23 leaves it in the low mantissa bits
    // (0x4B000000 + di), and dm_biased already subtracts that constant - no float->int conversion on the XU pipe
    const uint32_t dbits = __float_as_uint(di + 8388608.0f);
#if defined(MSP_EXPERIMENT_NO_GATHER)
    if (r.go) r.raw = 60000u;                                   // upper bound experiment: no memory access
#elif defined(MSP_EXPERIMENT_HOT_GATHER)
    if (r.go) r.raw = ldg_u16_nc(reinterpret_cast<const uint16_t*>(dmb + 2ull * (0x4B000000u + (dbits & 4095u)))) | 0x8000u;
#else
    if (r.go) r.raw = ldg_u16_nc(reinterpret_cast<const uint16_t*>(dmb + 2ull * dbits));
#endif
    return r;
  };
  auto cmp32 = [&](const Lane32& r, float Bz, bool& vis, bool& unc) {
#if MSP_CULL_MAGIC_I2F
    const float d = (__uint_as_float(0x4B000000u | r.raw) - 8388608.0f) * p.fscale;
#else
    const float d = (float)r.raw * p.fscale;
#endif
    const float mz = fmaf(3.6e-7f, d, Bz);  // 6u d + B_Z: d^ is within 2.1u d of raw*0.001, plus the rounding of z^ - d^
    const float t = r.zf - d;
    vis = r.go && (t < -mz);
    unc = r.go && (fabsf(t) <= mz);
  };

  // frames in which the whole tile is certainly outside the image: their mask words are zero, nothing else to do
  if (tid == 0) {
    int k = 0;
    for (int fi = 0; fi < nf; ++fi)
      if (s_c[fi].tile_cand != 0) s_flist[k++] = (uint8_t)fi;
    s_nlist = k;
  }
  if (MODE == 0 && word_ok) {
    for (int fi = 0; fi < nf; ++fi)
      if (s_c[fi].tile_cand == 0) mask_ptr[(int64_t)fi * p.mask_stride] = 0u;
  }
  __syncthreads();
  const int nlist = s_nlist;

#if MSP_CULL_DEFER_EXACT
  uint32_t dq = 0;  // warp-uniform fill of s_dq[warp]
  // one deferred point on this lane: exact re-evaluation; a visible point is OR-ed into its (already stored) mask word and
  // added to the count histogram.  Callers order the plain mask stores before this with __syncwarp().
  auto exact_deferred = [&](uint32_t e) {
    const int fi2 = (int)(e >> 8), j = (int)((e >> 5) & 7u), sl = (int)(e & 31u);
    const int idx = MSP_CULL_GROUP(warp, j) * 32 + sl;
    const int64_t f2 = f0 + fi2;
    const double* w = s_w + fi2 * 16;
    const bool fstd = s_c[fi2].std_frame != 0;
    double x = 0.0, y = 0.0, z = 0.0;
    if (idx < cnt) PtTraits<DT>::load(p.pts, base + idx, x, y, z);
    const Proj pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
    if (idx < cnt && in_image(pr, p.g) && pr.z > 0.0) {
      const int di = depth_index(pr, p.g);
      if (depth_test(pr.z, ldg_u16_nc(p.depth + f2 * p.depth_stride + di), p.g)) {
        atomicOr(p.mask + f2 * p.mask_stride + (int64_t)blockIdx.x * (kCullTilePoints / 32) + MSP_CULL_GROUP(warp, j), 1u << sl);
        if (COUNT) {
          const uint32_t id = (uint32_t)__ldg(p.inst + base + idx);
          if (id != 0u && id <= (uint32_t)p.n_objects) atomicAdd(&s_cnt[fi2 * kCntWords + (id >> 1)], 1u << ((id & 1u) * 16u));
        }
      }
    }
  };
#endif
  for (int li = 0; li < nlist; ++li) {
    const int fi = s_flist[li];
    const CullFrameConsts& c = s_c[fi];
    const int64_t f = f0 + fi;
    uint32_t myword = 0;  // lane j (< 8) ends up holding mask word j of this warp
    uint32_t nq = 0;      // lanes queued for the exact path

    // ---- A: group boxes against the four half-spaces (lane = form q x group g); skipped when the whole tile is out
    const float B = c.B;
    uint32_t gm = 0u;
    {
      const float4 a = *reinterpret_cast<const float4*>(c.box[lane >> 3]);
#if MSP_CULL_BOX_SMEM && MSP_CULL_BOX_ADDR
      float4 bc, bh;
      asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bc.x), "=f"(bc.y), "=f"(bc.z), "=f"(bc.w) : "r"(s_box_lane));
      asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(bh.x), "=f"(bh.y), "=f"(bh.z), "=f"(bh.w) : "r"(s_box_lane));
#elif MSP_CULL_BOX_SMEM
      const float4 bc = *reinterpret_cast<const float4*>(&s_box[warp][lane & 7][0]);
      const float4 bh = *reinterpret_cast<const float4*>(&s_box[warp][lane & 7][4]);
#endif
#if MSP_CULL_BOX_SMEM
      const float v = fmaf(a.x, bc.x, fmaf(a.y, bc.y, fmaf(a.z, bc.z, a.w)));
      const float r = __fmaf_ru(fabsf(a.x), bh.x, __fmaf_ru(fabsf(a.y), bh.y, __fmul_ru(fabsf(a.z), bh.z)));
#else
      const float v = fmaf(a.x, bcx, fmaf(a.y, bcy, fmaf(a.z, bcz, a.w)));
      const float r = __fmaf_ru(fabsf(a.x), bhx, __fmaf_ru(fabsf(a.y), bhy, __fmul_ru(fabsf(a.z), bhz)));
#endif
      const uint32_t gb = __ballot_sync(0xffffffffu, __fadd_ru(v, r) >= -B);
      gm = gb & (gb >> 8) & (gb >> 16) & (gb >> 24) & 0xffu;
    }

    if (gm != 0u) {
      const float4 a0 = *reinterpret_cast<const float4*>(c.box[0]);
      const float4 a2 = *reinterpret_cast<const float4*>(c.box[2]);
      const float4 az = *reinterpret_cast<const float4*>(c.fz);
      const float B2 = B + B, ct1 = c.ct1, half_ct0 = 0.5f - c.ct0, Bz = c.Bz;
      unsigned long long dm = c.dm_biased;
      asm volatile("" : "+l"(dm));  // keep the frame's depth base in registers (it is otherwise rebuilt per gather)
      if (MODE == 1) {
        wz = s_w + fi * 16 + 8;
        zrow = reinterpret_cast<unsigned int*>(p.zbuf + f * p.depth_stride);
      }
      if (p.stats != nullptr) st_f32 += __popc(gm);
#if MSP_CULL_PIPELINE
      // ---- B + C, software-pipelined inside the frame: iteration k projects pair k and issues its two depth gathers, then
      // finishes pair k-1 (depth comparison, ballots, queueing), whose gathers have had a whole projection to land.
      // A single left-over group travels as a pair whose second slot is empty (j = -1: no lane owns its word).
      Lane32 q0, q1;
      int qj0 = -1, qj1 = -1;
      bool have = false;
      q0.zf = q1.zf = 0.f, q0.raw = q1.raw = 0u, q0.go = q1.go = q0.unc = q1.unc = false;
      for (;;) {
        const bool more = gm != 0u;  // warp-uniform
        Lane32 r0, r1;
        int j0 = -1, j1 = -1;
        r0.zf = r1.zf = 0.f, r0.raw = r1.raw = 0u, r0.go = r1.go = r0.unc = r1.unc = false;
        if (more) {
          j0 = __ffs(gm) - 1;
          gm &= gm - 1u;
          r0 = eval32(j0, a0, a2, az, B, B2, ct1, half_ct0, dm);
          if (gm != 0u) {
            j1 = __ffs(gm) - 1;
            gm &= gm - 1u;
            r1 = eval32(j1, a0, a2, az, B, B2, ct1, half_ct0, dm);
          }
        }
        if (have) {
          bool v0, u0, v1, u1;
          cmp32(q0, Bz, v0, u0);
          cmp32(q1, Bz, v1, u1);
          const uint32_t bv0 = __ballot_sync(0xffffffffu, v0), bv1 = __ballot_sync(0xffffffffu, v1);
          if (lane == qj0) myword = bv0;
          if (lane == qj1) myword = bv1;
          const uint32_t bu0 = __ballot_sync(0xffffffffu, q0.unc | u0), bu1 = __ballot_sync(0xffffffffu, q1.unc | u1);
          if ((bu0 | bu1) != 0u) {  // rare, warp-uniform
            if (q0.unc | u0) s_q[warp][nq + __popc(bu0 & lt)] = (uint8_t)((qj0 << 5) | lane);
            nq += __popc(bu0);
            if (q1.unc | u1) s_q[warp][nq + __popc(bu1 & lt)] = (uint8_t)((qj1 << 5) | lane);
            nq += __popc(bu1);
          }
        }
        if (!more) break;
        q0 = r0, q1 = r1, qj0 = j0, qj1 = j1, have = true;
      }
#else
      // ---- B + C: candidate groups two at a time (two independent dependency chains per lane), then a leftover
      while (gm & (gm - 1u)) {
        const int j0 = __ffs(gm) - 1;
        gm &= gm - 1u;
        const int j1 = __ffs(gm) - 1;
        gm &= gm - 1u;
        const Lane32 r0 = eval32(j0, a0, a2, az, B, B2, ct1, half_ct0, dm);
        const Lane32 r1 = eval32(j1, a0, a2, az, B, B2, ct1, half_ct0, dm);
        bool v0, u0, v1, u1;
        cmp32(r0, Bz, v0, u0);
        cmp32(r1, Bz, v1, u1);
        const uint32_t bv0 = __ballot_sync(0xffffffffu, v0), bv1 = __ballot_sync(0xffffffffu, v1);
        if (lane == j0) myword = bv0;
        if (lane == j1) myword = bv1;
        const uint32_t bu0 = __ballot_sync(0xffffffffu, r0.unc | u0), bu1 = __ballot_sync(0xffffffffu, r1.unc | u1);
        if ((bu0 | bu1) != 0u) {  // rare, warp-uniform
          if (r0.unc | u0) s_q[warp][nq + __popc(bu0 & lt)] = (uint8_t)((j0 << 5) | lane);
          nq += __popc(bu0);
          if (r1.unc | u1) s_q[warp][nq + __popc(bu1 & lt)] = (uint8_t)((j1 << 5) | lane);
          nq += __popc(bu1);
        }
      }
      if (gm != 0u) {
        const int j0 = __ffs(gm) - 1;
        const Lane32 r0 = eval32(j0, a0, a2, az, B, B2, ct1, half_ct0, dm);
        bool v0, u0;
        cmp32(r0, Bz, v0, u0);
        const uint32_t bv0 = __ballot_sync(0xffffffffu, v0);
        if (lane == j0) myword = bv0;
        const uint32_t bu0 = __ballot_sync(0xffffffffu, r0.unc | u0);
        if (bu0 != 0u) {
          if (r0.unc | u0) s_q[warp][nq + __popc(bu0 & lt)] = (uint8_t)((j0 << 5) | lane);
          nq += __popc(bu0);
        }
      }
#endif
      if (p.stats != nullptr) st_exact += (lane == 0) ? nq : 0;
    }

    // ---- D: exact float64 reference-order evaluation of the queued lanes (rare)
#if MSP_CULL_DEFER_EXACT
    const bool defer = (MODE == 0) && nq != 0u && nq <= 32u;  // warp-uniform; larger bursts (non-standard frames) stay immediate
    if (nq != 0u && !defer) {
#else
    if (nq != 0u) {
#endif
      __syncwarp();
      const double* w = s_w + fi * 16;
      const bool fstd = c.std_frame != 0;
      for (uint32_t t0 = 0; t0 < nq; t0 += 32) {
        const uint32_t t = t0 + lane;
        if (t < nq) {
          const uint32_t code = s_q[warp][t];
          const int j = code >> 5, sl = code & 31;
          const int idx = MSP_CULL_GROUP(warp, j) * 32 + sl;
          double x = 0.0, y = 0.0, z = 0.0;
          if (idx < cnt) PtTraits<DT>::load(p.pts, base + idx, x, y, z);
          const Proj pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
          if (idx < cnt && in_image(pr, p.g) && pr.z > 0.0) {
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
      __syncwarp();
      if (MODE == 0 && lane < kCullPtsPerLane) {
        myword |= s_words[warp][lane];
        s_words[warp][lane] = 0u;
      }
      __syncwarp();
    }
    if (MODE == 0) {
      if (word_ok) mask_ptr[(int64_t)fi * p.mask_stride] = myword;
      if (COUNT) {
        // lanes 0..7 hold the final words; uniform groups: one shared-memory add per non-empty word
        const bool mine_uni = (lane < kCullPtsPerLane) && ((uni_mask >> lane) & 1u);
        if (mine_uni && myword != 0u && gid != 0u)
          atomicAdd(&s_cnt[fi * kCntWords + (gid >> 1)], (uint32_t)__popc(myword) << ((gid & 1u) * 16u));
        uint32_t mixed = __ballot_sync(0xffffffffu, (lane < kCullPtsPerLane) && !mine_uni && myword != 0u);
        while (mixed) {  // groups that straddle an object boundary (rare): aggregate by id within the warp
          const int j = __ffs(mixed) - 1;
          mixed &= mixed - 1;
          const uint32_t wv = __shfl_sync(0xffffffffu, myword, j);
          const uint32_t pack = (j < 2) ? ids_lo : (j < 4) ? ids_hi : (j < 6) ? ids_lo2 : ids_hi2;
          const uint32_t id = (pack >> ((uint32_t)(j & 1) * 16u)) & 0xffffu;
          const bool act = ((wv >> lane) & 1u) && id != 0u;
          const uint32_t am = __ballot_sync(0xffffffffu, act);
          if (act) {
            const uint32_t grp = __match_any_sync(am, id);
            if (lane == __ffs(grp) - 1) atomicAdd(&s_cnt[fi * kCntWords + (id >> 1)], (uint32_t)__popc(grp) << ((id & 1u) * 16u));
          }
        }
      }
    }
#if MSP_CULL_DEFER_EXACT
    if (defer) {  // after this frame's mask words have been stored
      __syncwarp();
      if (lane < nq) s_dq[warp][dq + lane] = (uint16_t)(((uint32_t)fi << 8) | s_q[warp][lane]);
      dq += nq;
      __syncwarp();  // s_q may be refilled by the next frame; queue entries and the plain mask stores of lanes 0..7 are visible
      if (dq >= 32u) {
        exact_deferred(s_dq[warp][dq - 32u + lane]);
        dq -= 32u;
        __syncwarp();
      }
    }
#endif
  }
#if MSP_CULL_DEFER_EXACT
  if (dq != 0u) {
    __syncwarp();
    if (lane < dq) exact_deferred(s_dq[warp][lane]);
  }
#endif
  if (COUNT) {
    __syncthreads();
    const int o = tid & 63;  // 64 ids per frame row, 4 frame rows per pass
    if (o < p.n_objects) {
      for (int fr = tid >> 6; fr < nf; fr += kCullThreads / 64) {
        const uint32_t v = (s_cnt[fr * kCntWords + ((o + 1) >> 1)] >> (((o + 1) & 1) * 16)) & 0xffffu;
        if (v != 0u) atomicAdd(p.counts + (int64_t)(f0 + fr) * p.n_objects + o, v);
      }
    }
  }
  if (p.stats != nullptr) {
    atomicAdd(p.stats + 0, (unsigned long long)st_f32);
    atomicAdd(p.stats + 1, (unsigned long long)st_exact);
  }
}

}  // namespace msp
