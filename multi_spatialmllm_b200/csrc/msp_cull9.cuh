// K1 fast path, second generation ("v9"): msp_project_visibility / msp_project_visibility_objects with standard intrinsics.
//
// Same arithmetic contract and the same float32-decision-with-proven-bounds scheme as msp_cull.cuh (read its header comment
// first: forms I0, Q1, I1, Q3, Z re-centred on the tile centre, bounds B / B_Z, pixel guard band ct1 * rcp(Z) + ct0, depth
// guard 6u d + B_Z, everything uncertain goes to the exact float64 reference-order code).  What changed is the shape of the
// work, driven by the round-1 ncu capture (353 M warp instructions per 2 000-frame launch, 60 % issue utilisation: the kernel
// is instruction-issue bound, not HBM bound):
//
//   * k_cull_prepare (one launch per call, ~2 % of K1's time) does everything that only depends on the POINTS once instead
//     of once per (tile, frame chunk): tile centre, re-centred float32 coordinates, super-group boxes, extents and the
//     per-tile instance-id table.  Its output (CullTilePrep, 13.6 KB per 1 024-point tile) is what the main kernel stages
//     into shared memory with ONE bulk async copy (TMA engine, UBLKCP) — no reductions, no re-centring in the hot kernel.
//   * the unit of evaluation is a SUPER-GROUP of 64 consecutive points, two per lane, stored as float2 (lo, hi): the whole
//     float32 projection runs on Blackwell's packed FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...), which take one issue slot
//     for two points (measured: profiles/r2/ubench_f32x2.txt — half the FMA-pipe rate, but the pipe was only 23 % busy).
//     A warp-uniform coefficient is a scalar-broadcast operand of the packed instruction (`R.F32` in SASS): one register.
//   * group classification runs with lane = FRAME: one pass tests a super-group box against the four half-spaces of 32
//     frames at once (no ballot, no per-frame head), leaves a 4-bit candidate mask per (warp, frame) in a register, and the
//     lanes write the all-zero mask words of their non-candidate (frame, super-group)s directly.  The main loop then only
//     visits frames that have a candidate (ffs over a ballot), in ascending frame order (depth-map locality in L1/L2).
//   * within a frame ALL candidate super-groups of the warp are projected and their depth gathers issued (up to eight
//     16-bit loads in flight per warp) before the first one is compared: the gather latency is paid once per frame, not
//     once per super-group (the kernel is bound by that latency at 32 warps x 64 registers per SM, not by bytes); lanes
//     0..3 collect the frame's eight mask words and store them once per frame;
//   * uncertain lanes are queued per warp ACROSS frames and the exact float64 path runs on 32 of them at a time, OR-ing
//     visible bits into the stored words (the validated round-1 "deferred exact" variant, now the only exact path — a
//     non-standard frame simply queues every lane and becomes the all-float64 kernel at full lane utilisation);
//   * fused per-(frame, object) counting works for ANY number of objects: a tile of 1 024 consecutive points touches few
//     instances, so the prepare pass gives each tile a 15-entry table of its ids and marks every super-group whose points
//     share one id; such a super-group adds its popcount to counts[frame][id] with ONE global RED per frame (no shared
//     histogram, no end-of-CTA barrier), mixed ones aggregate by id within the warp first; ids that do not fit the table
//     (> 15 distinct in one tile, e.g. shuffled clouds) are looked up in the caller's id array.
//
// Bit-identical to k_frame_loop (all-float64) and to the round-1 kernel by construction: a float32 decision is only taken
// when the bounds prove it; the packed operations are IEEE single roundings per half in the same association order as the
// scalar code (fl(a0 x' + fl(a1 y' + fl(a2 z' + d')))), the pixel value I0 * rcp(Z) is now used as the EXACT product
// inside an FMA (one rounding fewer than the bound allows for).
#pragma once

#include <stddef.h>

#include "msp_cull.cuh"
#include "msp_device.cuh"

namespace msp {

#ifndef MSP_C9_THREADS
#define MSP_C9_THREADS 128     // 4 warps, 1 024-point tiles, 8 CTAs/SM: same warps per SM as 256 x 4 but finer-grained CTAs (0.373 -> 0.362 ms)
#endif
constexpr int kC9Threads = MSP_C9_THREADS;
constexpr int kC9Warps = kC9Threads / 32;
constexpr int kC9SgPerWarp = 4;                                   // super-groups (64 points, two per lane) per warp
constexpr int kC9TileSgs = kC9Warps * kC9SgPerWarp;               // 16
constexpr int kC9TilePoints = kC9TileSgs * 64;                    // 1024
constexpr int kC9TileWords = kC9TilePoints / 32;                  // 32 mask words per tile and frame
#ifndef MSP_C9_FRAMES
#define MSP_C9_FRAMES 32
#endif
constexpr int kC9Frames = MSP_C9_FRAMES;                          // frames per CTA, a multiple of 32
constexpr int kC9Rounds = kC9Frames / 32;
static_assert(kC9Frames % 32 == 0 && kC9Frames <= 256, "frames per CTA");
constexpr int kC9Slots = 16;                                      // local instance-id slots per tile (slot 15 is never used: 15 ids)
#ifndef MSP_C9_MINBLOCKS
#define MSP_C9_MINBLOCKS (1024 / MSP_C9_THREADS)   // 32 warps per SM at 64 registers
#endif
#ifndef MSP_C9_STRIDED
#define MSP_C9_STRIDED 1       // 1: warp w owns super-groups w, w+W, w+2W, w+3W for W warps per CTA (even work per warp), 0: 4w .. 4w+3
#endif
#ifndef MSP_C9_COEF_PER_SG
#define MSP_C9_COEF_PER_SG 1   // 1: the 12 frame coefficients are re-read from shared memory for every super-group (fewer live registers
                               //    next to the up to 16 registers of gather state; 0 measured the same within 0.5 %)
#endif
constexpr uint32_t kC9DeferCap = 288;                             // < 32 left over + at most 4 x 64 appended per frame
#if MSP_C9_STRIDED
#define MSP_C9_SG(warp, k) ((warp) + kC9Warps * (k))
#else
#define MSP_C9_SG(warp, k) ((warp) * kC9SgPerWarp + (k))
#endif

constexpr uint8_t kC9SlotNone = 0xFE;      // no object (id 0, id out of range, slot beyond the cloud)
constexpr uint8_t kC9SlotOverflow = 0xFD;  // the id is not in the tile's table: look it up in inst[] and count with a global atomic
constexpr uint8_t kC9SlotMixed = 0xFF;     // (super-group only) its points carry different ids: per-lane path

struct __align__(16) CullTilePrep {
  float2 pts[kC9TileSgs][3][32];  // [sg][x|y|z][lane] = (point 64 sg + lane, point 64 sg + 32 + lane) re-centred; NaN = empty slot
  float4 box[kC9TileSgs][2];      // [sg] (cx,cy,cz,-), (hx,hy,hz,-): box of the super-group's re-centred points, half sizes rounded up
  float centre[3];                // tile centre (float32, any finite value is correct)
  float ext[3];                   // max |x'|, |y'|, |z'| over the tile
  float pad0[2];
  uint16_t slot_id[kC9Slots];     // local slot -> instance id (0 = unused)
  uint8_t sg_slot[kC9TileSgs];    // 0..14 all points share the id in that slot, kC9SlotNone, or kC9SlotMixed
  uint8_t pt_slot[kC9TilePoints]; // per point: 0..14, kC9SlotNone or kC9SlotOverflow
};
static_assert(sizeof(CullTilePrep) % 16 == 0 && offsetof(CullTilePrep, pt_slot) % 16 == 0, "bulk copy sizes");

struct __align__(16) C9Frame {    // per (tile, frame); 112 bytes = 7 x 16 so that lane = frame LDS.128 accesses are conflict-free
  float a0[4];                    // I0 = sx (fx X + cx Z) as {x, y, z, d'} relative to the tile centre
  float a2[4];                    // I1
  float az[4];                    // Z
  float q1[4];                    // Q1 = dw Z - I0 (classification only)
  float q3[4];                    // Q3 = dh Z - I1
  float ct1, hct0, Bz, B;         // pixel guard band dt = ct1 rcp(Z) + ct0 as (ct1, 0.5 - ct0); depth-form bound; bound of the in-image forms
  unsigned long long dm_biased;   // depth map byte address - 2 * 0x4B000000 (gather index = as_uint(di + 2^23))
  int std_frame;                  // bottom row of w2c is exactly [0,0,0,1]
  uint32_t pad;
};
static_assert(sizeof(C9Frame) == 112, "C9Frame layout");

// ---- packed float32x2 arithmetic (one issue slot for two IEEE single operations; FFMA2 / FADD2 / FMUL2 in SASS) ----------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
// (c, c): ptxas recognises the repeated register and emits the packed instruction's scalar-broadcast operand form (`R6.F32` in SASS),
// so a warp-uniform coefficient costs ONE register, not a duplicated pair
__device__ __forceinline__ f32x2 bc2(float c) { return pk2(c, c); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 lds64(uint32_t addr) {
  f32x2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, f32x2& a, f32x2& b) {
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

// ================================================================================================================
// prepare pass: everything that depends on the points only.  One CTA per 1 024-point tile, warp w owns super-groups 4w..4w+3.
// ================================================================================================================
template <int DT>
__global__ void __launch_bounds__(kC9Threads) k_cull_prepare(const void* __restrict__ pts, int64_t n,
                                                             const uint16_t* __restrict__ inst, int32_t n_objects,
                                                             CullTilePrep* __restrict__ out) {
  __shared__ int s_red[9];           // ordered-int min xyz, max xyz, then uint max |x'|,|y'|,|z'|
  __shared__ uint32_t s_ids[kC9Slots];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kC9TilePoints;
  const int cnt = (int)min((int64_t)kC9TilePoints, n - base);
  CullTilePrep& o = out[blockIdx.x];
  const unsigned char* gp = static_cast<const unsigned char*>(pts) + base * PtTraits<DT>::kBytes;
  if (tid == 0) {
    s_red[0] = s_red[1] = s_red[2] = 0x7fffffff;
    s_red[3] = s_red[4] = s_red[5] = (int)0x80000000;
    s_red[6] = s_red[7] = s_red[8] = 0;
  }
  if (tid < kC9Slots) s_ids[tid] = 0u;
  __syncthreads();

  const float qnan = __int_as_float(0x7fc00000), finf = __int_as_float(0x7f800000);
  float X[kC9SgPerWarp][2], Y[kC9SgPerWarp][2], Z[kC9SgPerWarp][2];
  {
    float lx = finf, ly = finf, lz = finf, hx = -finf, hy = -finf, hz = -finf;
#pragma unroll
    for (int k = 0; k < kC9SgPerWarp; ++k)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = (warp * kC9SgPerWarp + k) * 64 + 32 * h + lane;
        X[k][h] = Y[k][h] = Z[k][h] = qnan;  // slots beyond the cloud and non-finite points: every comparison on them is false
        if (idx < cnt) {
          float x, y, z;
          load_point_f32<DT>(gp, idx, x, y, z);
          if (isfinite(x) && isfinite(y) && isfinite(z)) {
            X[k][h] = x, Y[k][h] = y, Z[k][h] = z;
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
  float tcx = 0.5f * (ord2f(s_red[0]) + ord2f(s_red[3])), tcy = 0.5f * (ord2f(s_red[1]) + ord2f(s_red[4])),
        tcz = 0.5f * (ord2f(s_red[2]) + ord2f(s_red[5]));
  if (!isfinite(tcx) || !isfinite(tcy) || !isfinite(tcz)) tcx = tcy = tcz = 0.f;  // empty tile; any finite centre is correct

  float ex = 0.f, ey = 0.f, ez = 0.f;
#pragma unroll
  for (int k = 0; k < kC9SgPerWarp; ++k) {
    const int sg = warp * kC9SgPerWarp + k;
    float lx = finf, ly = finf, lz = finf, hx = -finf, hy = -finf, hz = -finf;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (DT == 2) {
        // float64 inputs: subtract the centre in float64 and round ONCE, so that the error of x' is u |x'| and not u |x|
        if (X[k][h] == X[k][h]) {
          const int idx = sg * 64 + 32 * h + lane;
          double x64, y64, z64;
          PtTraits<DT>::load(gp, idx, x64, y64, z64);
          X[k][h] = __double2float_rn(x64 - (double)tcx);
          Y[k][h] = __double2float_rn(y64 - (double)tcy);
          Z[k][h] = __double2float_rn(z64 - (double)tcz);
        }
      } else {
        X[k][h] -= tcx, Y[k][h] -= tcy, Z[k][h] -= tcz;  // NaN stays NaN
      }
      const float x = X[k][h], y = Y[k][h], z = Z[k][h];
      if (x == x) {
        ex = fmaxf(ex, fabsf(x)), ey = fmaxf(ey, fabsf(y)), ez = fmaxf(ez, fabsf(z));
        lx = fminf(lx, x), ly = fminf(ly, y), lz = fminf(lz, z);
        hx = fmaxf(hx, x), hy = fmaxf(hy, y), hz = fmaxf(hz, z);
      }
    }
    const float gx0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lx))), gx1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hx)));
    const float gy0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(ly))), gy1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hy)));
    const float gz0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lz))), gz1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hz)));
    if (lane == 0) {
      // centre / half size with the half size rounded up so that [c-h, c+h] certainly contains the super-group's x'
      const float k1 = 1.0f + 2.384185791015625e-07f, k2 = 4.76837158203125e-07f;  // 1+2^-22, 2^-21
      const float bcx = 0.5f * (gx0 + gx1), bcy = 0.5f * (gy0 + gy1), bcz = 0.5f * (gz0 + gz1);
      const float bhx = __fmaf_ru(0.5f * (gx1 - gx0), k1, k2 * fmaxf(fabsf(gx0), fabsf(gx1)));
      const float bhy = __fmaf_ru(0.5f * (gy1 - gy0), k1, k2 * fmaxf(fabsf(gy0), fabsf(gy1)));
      const float bhz = __fmaf_ru(0.5f * (gz1 - gz0), k1, k2 * fmaxf(fabsf(gz0), fabsf(gz1)));
      // an empty super-group has inf - inf = NaN in its centre: every test on it is false, it is never a candidate
      o.box[sg][0] = make_float4(bcx, bcy, bcz, 0.f);
      o.box[sg][1] = make_float4(bhx, bhy, bhz, 0.f);
    }
    o.pts[sg][0][lane] = make_float2(X[k][0], X[k][1]);
    o.pts[sg][1][lane] = make_float2(Y[k][0], Y[k][1]);
    o.pts[sg][2][lane] = make_float2(Z[k][0], Z[k][1]);
  }
  {
    const uint32_t ux = __reduce_max_sync(0xffffffffu, __float_as_uint(ex));
    const uint32_t uy = __reduce_max_sync(0xffffffffu, __float_as_uint(ey));
    const uint32_t uz = __reduce_max_sync(0xffffffffu, __float_as_uint(ez));
    if (lane == 0) atomicMax(&s_red[6], (int)ux), atomicMax(&s_red[7], (int)uy), atomicMax(&s_red[8], (int)uz);
  }

  // ---- instance ids: the tile's table of distinct ids (first come, first served) and every point's local slot
  uint32_t id[kC9SgPerWarp][2];
  if (inst != nullptr) {
#pragma unroll
    for (int k = 0; k < kC9SgPerWarp; ++k)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = (warp * kC9SgPerWarp + k) * 64 + 32 * h + lane;
        uint32_t v = (idx < cnt) ? (uint32_t)__ldg(inst + base + idx) : 0u;
        if (v > (uint32_t)n_objects) v = 0u;
        id[k][h] = v;
        const uint32_t peers = __match_any_sync(0xffffffffu, v);
        if (v != 0u && lane == __ffs(peers) - 1) {
          for (int s = 0; s < kC9Slots - 1; ++s) {
            const uint32_t old = atomicCAS(&s_ids[s], 0u, v);
            if (old == 0u || old == v) break;
          }
        }
      }
  }
  __syncthreads();
  if (tid == 0) {
    o.centre[0] = tcx, o.centre[1] = tcy, o.centre[2] = tcz;
    o.ext[0] = __int_as_float(s_red[6]), o.ext[1] = __int_as_float(s_red[7]), o.ext[2] = __int_as_float(s_red[8]);
    o.pad0[0] = o.pad0[1] = 0.f;
  }
  if (tid < kC9Slots) o.slot_id[tid] = (inst != nullptr) ? (uint16_t)s_ids[tid] : (uint16_t)0;
  if (inst != nullptr) {
#pragma unroll
    for (int k = 0; k < kC9SgPerWarp; ++k) {
      const int sg = warp * kC9SgPerWarp + k;
      uint32_t sl[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t v = id[k][h];
        uint32_t s = kC9SlotNone;
        if (v != 0u) {
          s = kC9SlotOverflow;
          for (int t = 0; t < kC9Slots - 1; ++t)
            if (s_ids[t] == v) s = (uint32_t)t;
        }
        sl[h] = s;
        o.pt_slot[sg * 64 + 32 * h + lane] = (uint8_t)s;
      }
      // uniform super-group: every point INSIDE the cloud carries the same id (slots beyond it never become visible)
      const uint32_t first = __shfl_sync(0xffffffffu, id[k][0], 0);
      const uint32_t first_slot = __shfl_sync(0xffffffffu, sl[0], 0);
      const int i0 = sg * 64 + lane, i1 = i0 + 32;
      const bool same = __all_sync(0xffffffffu, (id[k][0] == first || i0 >= cnt) && (id[k][1] == first || i1 >= cnt));
      if (lane == 0) o.sg_slot[sg] = (uint8_t)((same && first_slot != kC9SlotOverflow) ? first_slot : kC9SlotMixed);
    }
  } else {
    if (tid < kC9TileSgs) o.sg_slot[tid] = kC9SlotNone;
  }
}

// ================================================================================================================
// per-(tile, frame) float32 forms and rigorous error bounds, float64 arithmetic (identical to the round-1 kernel)
// ================================================================================================================
__device__ __forceinline__ void c9_frame_consts(const double* __restrict__ w, const Intrinsics& K, const ImageGeom& g,
                                                const float* __restrict__ centre, const float* __restrict__ extf,
                                                const uint16_t* depth_map, C9Frame& c) {
  const float finf = __int_as_float(0x7f800000);
  const double kUp = 1.0 + 2.4e-07;  // extents: cover the rounding of x' itself and of float64 inputs
  const double ext[3] = {(double)extf[0] * kUp, (double)extf[1] * kUp, (double)extf[2] * kUp};
  const double tc[3] = {(double)centre[0], (double)centre[1], (double)centre[2]};
  const double fx = K.k[0], cx = K.k[2], fy = K.k[5], cy = K.k[6];
  const double sx = g.sx, sy = g.sy, dw = (double)g.depth_w, dh = (double)g.depth_h;
  double A[5][4], M[5], mag[3];
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
  const double u = 5.9604644775390625e-08, tiny = 1.4210854715202004e-14;  // 2^-24, 2^-46
  double Bf[5];
  float co[5][4];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const double dprime = A[q][0] * tc[0] + A[q][1] * tc[1] + A[q][2] * tc[2] + A[q][3];
    const double Aext = fabs(A[q][0]) * ext[0] + fabs(A[q][1]) * ext[1] + fabs(A[q][2]) * ext[2];
    Bf[q] = 8.0 * u * (Aext + fabs(dprime)) + tiny * M[q];
    co[q][0] = __double2float_rn(A[q][0]), co[q][1] = __double2float_rn(A[q][1]), co[q][2] = __double2float_rn(A[q][2]);
    co[q][3] = __double2float_rn(dprime);
  }
  // Q1 and Q3 are *derived* in the point loop as fl(dw Z^ - I0^): their error is B_I0 + dw B_Z + one rounding
  const double bq1 = 1.01 * (Bf[0] + dw * Bf[4]) + 0.13 * Bf[1];
  const double bq3 = 1.01 * (Bf[2] + dh * Bf[4]) + 0.13 * Bf[3];
  const double B = fmax(fmax(Bf[0], Bf[2]), fmax(fmax(Bf[1], Bf[3]), fmax(bq1, bq3)));
  float Bf32 = __double2float_ru(B);
  const float ct1 = __double2float_ru(1.01 * fmax(Bf[0] + dw * Bf[4], Bf[2] + dh * Bf[4]));
  const float ct0 = __double2float_ru(8.0 * u * fmax(dw, dh) + 1e-9);
  const float Bz = __double2float_ru(Bf[4]);
  c.std_frame = (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
  if (!c.std_frame && isfinite(B)) {
    // perspective row present: the forms are not the image-plane tests — every point goes to the exact path
#pragma unroll
    for (int q = 0; q < 4; ++q) co[q][0] = co[q][1] = co[q][2] = co[q][3] = 0.f;
    Bf32 = finf;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c.a0[k] = co[0][k];
    c.a2[k] = co[2][k];
    c.az[k] = co[4][k];
    c.q1[k] = co[1][k];
    c.q3[k] = co[3][k];
  }
  c.ct1 = ct1;
  c.hct0 = 0.5f - ct0;
  c.Bz = Bz;
  c.B = Bf32;
  c.dm_biased = (unsigned long long)depth_map - 2ull * 0x4B000000ull;
  c.pad = 0u;
}

struct Cull9Params {
  CullParams p;               // the caller's arrays (msp_cull.cuh)
  const CullTilePrep* prep;   // one per tile (k_cull_prepare)
  float nfdw, nfdh, fdw;      // -depth_w, -depth_h, depth_w as float32 (exact)
  float fscale;               // float32 rounding of depth_scale (its error is inside the 6u d guard)
  float xclip, yclip;         // dw - 1 + 2^23, dh - 1
};

// ================================================================================================================
// main kernel
// ================================================================================================================
template <int DT, bool COUNT>
__global__ void __launch_bounds__(kC9Threads, MSP_C9_MINBLOCKS) k_frame_cull9(const Cull9Params q) {
  extern __shared__ __align__(128) unsigned char s_raw[];  // CullTilePrep (without pt_slot when !COUNT)
  __shared__ __align__(16) C9Frame s_c[kC9Frames];
  __shared__ uint16_t s_dq[kC9Warps][kC9DeferCap];              // (frame in chunk << 8) | (k << 6) | (half << 5) | lane
  __shared__ __align__(8) uint64_t s_bar;
  const CullParams& p = q.p;
  const CullTilePrep& T = *reinterpret_cast<const CullTilePrep*>(s_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kC9TilePoints;
  const int cnt = (int)min((int64_t)kC9TilePoints, p.n - base);
  const int f0 = blockIdx.y * kC9Frames;
  const int nf = min(kC9Frames, p.F - f0);
  const CullTilePrep* gprep = q.prep + blockIdx.x;

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    constexpr uint32_t bytes = COUNT ? (uint32_t)sizeof(CullTilePrep) : (uint32_t)offsetof(CullTilePrep, pt_slot);
    mbar_arrive_expect_tx(&s_bar, bytes);
    bulk_g2s(s_raw, gprep, bytes, &s_bar);
  }
  // per-frame constants while the tile is in flight (the header is read straight from global memory / L2)
  if (tid < nf) {
    const float centre[3] = {__ldg(&gprep->centre[0]), __ldg(&gprep->centre[1]), __ldg(&gprep->centre[2])};
    const float extf[3] = {__ldg(&gprep->ext[0]), __ldg(&gprep->ext[1]), __ldg(&gprep->ext[2])};
    c9_frame_consts(p.w2c + (int64_t)(f0 + tid) * 16, p.K, p.g, centre, extf, p.depth + (int64_t)(f0 + tid) * p.depth_stride, s_c[tid]);
  }
  mbar_wait(&s_bar, 0);
  __syncthreads();

  const uint32_t lt = (1u << lane) - 1u;
  const int64_t word0 = (int64_t)blockIdx.x * kC9TileWords;
  uint32_t okm = 0;  // bit k: the two mask words of this warp's super-group k exist (mask_stride is a multiple of 4)
#pragma unroll
  for (int k = 0; k < kC9SgPerWarp; ++k)
    if (word0 + 2 * MSP_C9_SG(warp, k) < p.mask_stride) okm |= 1u << k;
  uint32_t* mask0 = p.mask + (int64_t)f0 * p.mask_stride + word0;  // + fi * mask_stride + 2 * sg
  asm volatile("" : "+l"(mask0));  // kept in a register pair instead of being rebuilt from blockIdx at every store
  // this lane's points of super-group k are at shared address s_lane + kSgStride k (+0 x, +256 y, +512 z); the address is made
  // opaque so that it stays in a register instead of being re-derived from the thread id at every use
  uint32_t s_lane = smem_u32(s_raw) + (uint32_t)MSP_C9_SG(warp, 0) * 768u + (uint32_t)lane * 8u;
  asm volatile("" : "+r"(s_lane));
  constexpr uint32_t kSgStride = (uint32_t)(MSP_C9_SG(0, 1) - MSP_C9_SG(0, 0)) * 768u;
  // ---- A: classification, lane = frame.  gmr[r] bit k: super-group k of this warp may be visible in frame 32 r + lane.
  uint32_t gmr[kC9Rounds];
#pragma unroll
  for (int r = 0; r < kC9Rounds; ++r) {
    const int fi = 32 * r + lane;
    uint32_t gm = 0u;
    if (fi < nf) {
      const C9Frame& c = s_c[fi];
      const float4 a0 = *reinterpret_cast<const float4*>(c.a0), a2 = *reinterpret_cast<const float4*>(c.a2);
      const float4 q1 = *reinterpret_cast<const float4*>(c.q1), q3 = *reinterpret_cast<const float4*>(c.q3);
      const float nB = -c.B;
#pragma unroll
      for (int k = 0; k < kC9SgPerWarp; ++k) {
        const float4 bc = T.box[MSP_C9_SG(warp, k)][0], bh = T.box[MSP_C9_SG(warp, k)][1];
        // form value at the box centre + |a| . half size (rounded up) >= -B for all four forms; NaN (empty box, NaN pose) -> false
        const float v0 = fmaf(a0.x, bc.x, fmaf(a0.y, bc.y, fmaf(a0.z, bc.z, a0.w)));
        const float r0 = __fmaf_ru(fabsf(a0.x), bh.x, __fmaf_ru(fabsf(a0.y), bh.y, __fmul_ru(fabsf(a0.z), bh.z)));
        const float v1 = fmaf(q1.x, bc.x, fmaf(q1.y, bc.y, fmaf(q1.z, bc.z, q1.w)));
        const float r1 = __fmaf_ru(fabsf(q1.x), bh.x, __fmaf_ru(fabsf(q1.y), bh.y, __fmul_ru(fabsf(q1.z), bh.z)));
        const float v2 = fmaf(a2.x, bc.x, fmaf(a2.y, bc.y, fmaf(a2.z, bc.z, a2.w)));
        const float r2 = __fmaf_ru(fabsf(a2.x), bh.x, __fmaf_ru(fabsf(a2.y), bh.y, __fmul_ru(fabsf(a2.z), bh.z)));
        const float v3 = fmaf(q3.x, bc.x, fmaf(q3.y, bc.y, fmaf(q3.z, bc.z, q3.w)));
        const float r3 = __fmaf_ru(fabsf(q3.x), bh.x, __fmaf_ru(fabsf(q3.y), bh.y, __fmul_ru(fabsf(q3.z), bh.z)));
        const bool cand = (__fadd_ru(v0, r0) >= nB) && (__fadd_ru(v1, r1) >= nB) && (__fadd_ru(v2, r2) >= nB) && (__fadd_ru(v3, r3) >= nB);
        if (cand) gm |= 1u << k;
      }
      // a frame without candidates never reaches the main loop: its (zero) mask words are written here, by the frame's lane
      // (in a frame WITH candidates lanes 0..3 of the main loop write all eight words, zero for the non-candidate super-groups)
      uint32_t* zrow = mask0 + (int64_t)fi * p.mask_stride;
#pragma unroll
      for (int k = 0; k < kC9SgPerWarp; ++k)
        if (gm == 0u && ((okm >> k) & 1u))
          *reinterpret_cast<uint2*>(zrow + 2 * MSP_C9_SG(warp, k)) = make_uint2(0u, 0u);
    }
    gmr[r] = gm;
  }

  uint32_t dq = 0;  // warp-uniform fill of s_dq[warp]
  // one deferred point on this lane: exact re-evaluation in the reference's float64 operation order; a visible point is OR-ed
  // into its (already stored) mask word and counted.  Callers order the plain mask stores before this with __syncwarp().
  auto exact_deferred = [&](uint32_t e) {
    const int fi = (int)(e >> 8), k = (int)((e >> 6) & 3u), sl = (int)(e & 63u);  // sl = half * 32 + lane
    const int sg = MSP_C9_SG(warp, k);
    const int idx = sg * 64 + sl;
    const int64_t f = f0 + fi;
    const double* w = p.w2c + f * 16;
    const bool fstd = s_c[fi].std_frame != 0;
    double x = 0.0, y = 0.0, z = 0.0;
    if (idx < cnt) PtTraits<DT>::load(p.pts, base + idx, x, y, z);
    const Proj pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
    if (idx < cnt && in_image(pr, p.g) && pr.z > 0.0) {
      const int di = depth_index(pr, p.g);
      if (depth_test(pr.z, ldg_u16_nc(p.depth + f * p.depth_stride + di), p.g)) {
        atomicOr(p.mask + f * p.mask_stride + word0 + 2 * sg + (sl >> 5), 1u << (sl & 31));
        if (COUNT) {
          const uint32_t s = T.pt_slot[idx];
          if (s < (uint32_t)(kC9Slots - 1)) atomicAdd(p.counts + f * p.n_objects + (int)T.slot_id[s] - 1, 1u);
          else if (s == kC9SlotOverflow) atomicAdd(p.counts + f * p.n_objects + (int)__ldg(p.inst + base + idx) - 1, 1u);
        }
      }
    }
  };
  // run the exact path on the queue's tail, 32 entries at a time, while it holds at least `keep_below` entries
  auto flush = [&](uint32_t keep_below) {
    __syncwarp();  // queue entries and the plain mask stores of lane 0 are visible to the whole warp
    while (dq >= keep_below && dq != 0u) {
      const uint32_t nb = dq < 32u ? dq : 32u;
      if (lane < nb) exact_deferred(s_dq[warp][dq - nb + lane]);
      dq -= nb;
    }
    __syncwarp();  // the queue may be refilled
  };

  // ---- B + C: frames with candidates, ascending.  Per frame the warp projects ALL its candidate super-groups and issues
  // their depth gathers (up to eight loads in flight per warp), looks the next candidate frame up, and only then compares
  // them: the wait is paid once per frame.  Nothing is pending at the loop's back-edge, and the only wait is for "all gathers
  // of this frame": ptxas puts every gather on ONE scoreboard, so a deeper pipeline (a second frame or super-group in flight
  // while the first is compared) waits for the younger loads as well (profiles/r2/k1_tuning.md, v10 / v17 / v18).
  // A lane that does not gather (outside the image, or uncertain) reads pixel 0 of the map (always mapped, L1-resident) and
  // carries z = +inf ("certainly not visible") or z = NaN ("uncertain"): t = z - d is then +inf / NaN for any d, so
  // `t < -mz` is false and the uncertainty test !(|t| > mz) is false / true as required — the comparison needs no predicates.
  const float finf = __int_as_float(0x7f800000), qnan = __int_as_float(0x7fc00000);
  const int my_sg = MSP_C9_SG(warp, lane & 3);  // lanes 0..3 own the two mask words (and the counter slot) of super-group k = lane
  const bool my_store = (lane < kC9SgPerWarp) && ((okm >> lane) & 1u);
  uint32_t my_slot = kC9SlotNone;  // COUNT: sg_slot of super-group k = lane (lanes 0..3)
  if (COUNT && lane < kC9SgPerWarp) my_slot = T.sg_slot[my_sg];
  // lanes 0..3: instance id - 1 of the lane's super-group when all its points share one id, else -1
  int my_obj = -1;
  if (COUNT && my_slot < (uint32_t)(kC9Slots - 1)) my_obj = (int)T.slot_id[my_slot] - 1;
  const bool any_mixed = COUNT && __any_sync(0xffffffffu, my_slot == kC9SlotMixed);
  // lanes 0..3 store their super-group's two words of frame fi (zero for a non-candidate super-group) and count them
  auto store_frame = [&](int fi, uint32_t pwl, uint32_t pwh) {
    const int64_t f = f0 + fi;
    if (my_store) *reinterpret_cast<uint2*>(p.mask + f * p.mask_stride + word0 + 2 * my_sg) = make_uint2(pwl, pwh);
    if (COUNT) {
      const bool nz = (pwl | pwh) != 0u;  // only lanes 0..3 can hold a non-zero word
      if (nz && my_obj >= 0) atomicAdd(p.counts + f * p.n_objects + my_obj, (uint32_t)(__popc(pwl) + __popc(pwh)));
      if (any_mixed) {
        // super-groups that straddle an object boundary (rare on mesh-ordered clouds): aggregate by slot within the warp
        uint32_t mg = __ballot_sync(0xffffffffu, nz && my_slot == kC9SlotMixed);
#pragma unroll 1
        while (mg != 0u) {
          const int k = __ffs(mg) - 1;
          mg &= mg - 1u;
          const uint32_t wvl = __shfl_sync(0xffffffffu, pwl, k), wvh = __shfl_sync(0xffffffffu, pwh, k);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int idx = MSP_C9_SG(warp, k) * 64 + 32 * h + lane;
            const uint32_t sl = T.pt_slot[idx];
            const bool bit = (((h ? wvh : wvl) >> lane) & 1u) != 0u;
            const bool act = bit && sl < (uint32_t)(kC9Slots - 1);
            const uint32_t am = __ballot_sync(0xffffffffu, act);
            if (act) {
              const uint32_t grp = __match_any_sync(am, sl);
              if (lane == __ffs(grp) - 1) atomicAdd(p.counts + f * p.n_objects + (int)T.slot_id[sl] - 1, (uint32_t)__popc(grp));
            }
            if (bit && sl == kC9SlotOverflow) atomicAdd(p.counts + f * p.n_objects + (int)__ldg(p.inst + base + idx) - 1, 1u);
          }
        }
      }
    }
    if (dq >= 32u) flush(32u);  // every queued point belongs to a frame whose words have been stored
  };
#pragma unroll 1
  for (int r = 0; r < kC9Rounds; ++r) {
    uint32_t fm = __ballot_sync(0xffffffffu, gmr[r] != 0u);
    int fl_next = __ffs(fm) - 1;                                              // -1: no candidate frame in this round
    uint32_t g_next = __shfl_sync(0xffffffffu, gmr[r], fl_next & 31);
    fm &= fm - 1u;
#pragma unroll 1
    while (fl_next >= 0) {
      const uint32_t g = g_next;
      const int fi = 32 * r + fl_next;
      const uint32_t s_cf = smem_u32(&s_c[fi]);
#if !MSP_C9_COEF_PER_SG
      const float4 a0 = lds_f4(s_cf + 0), a2 = lds_f4(s_cf + 16), az = lds_f4(s_cf + 32);
#endif
      const float4 sc4 = lds_f4(s_cf + 80);  // ct1, 1/2 - ct0, B_Z, B
      const f32x2 ct1_2 = bc2(sc4.x), hct0_2 = bc2(sc4.y);
      const float B2q = sc4.w + sc4.w, nB = -sc4.w;
      const unsigned long long dm = lds64(s_cf + 96);
      // the frame's candidate super-groups in flight: float32 z of the two points of a lane, the two gathered raw depths
      f32x2 bz[kC9SgPerWarp];
      uint32_t brl[kC9SgPerWarp], brh[kC9SgPerWarp];  // statically indexed: registers
#pragma unroll
      for (int k = 0; k < kC9SgPerWarp; ++k) {
        if (!((g >> k) & 1u)) continue;  // warp-uniform
        const uint32_t pa = s_lane + kSgStride * (uint32_t)k;
        const f32x2 xx = lds64(pa), yy = lds64(pa + 256u), zz = lds64(pa + 512u);
#if MSP_C9_COEF_PER_SG
        const float4 a0 = lds_f4(s_cf + 0), a2 = lds_f4(s_cf + 16), az = lds_f4(s_cf + 32);
#endif
        const f32x2 I0 = fma2(bc2(a0.x), xx, fma2(bc2(a0.y), yy, fma2(bc2(a0.z), zz, bc2(a0.w))));
        const f32x2 I1 = fma2(bc2(a2.x), xx, fma2(bc2(a2.y), yy, fma2(bc2(a2.z), zz, bc2(a2.w))));
        const f32x2 Zp = fma2(bc2(az.x), xx, fma2(bc2(az.y), yy, fma2(bc2(az.z), zz, bc2(az.w))));
        const f32x2 nQ1 = fma2(bc2(q.nfdw), Zp, I0);  // I0 - dw Z = -Q1 (same rounding as fl(dw Z - I0))
        const f32x2 nQ3 = fma2(bc2(q.nfdh), Zp, I1);
        float I0l, I0h, I1l, I1h, Zl, Zh, nq1l, nq1h, nq3l, nq3h;
        upk2(I0, I0l, I0h), upk2(I1, I1l, I1h), upk2(Zp, Zl, Zh), upk2(nQ1, nq1l, nq1h), upk2(nQ3, nq3l, nq3h);
        // m = min(I0, Q1, I1, Q3); NaN only for empty / non-finite slots: never visible, never queued
        const float ml = fminf(fminf(I0l, I1l), -fmaxf(nq1l, nq3l));
        const float mh = fminf(fminf(I0h, I1h), -fmaxf(nq1h, nq3h));
        const f32x2 s = pk2(rcp_approx(-Zl), rcp_approx(-Zh));  // s = -1/Z (max relative error 2^-23, covered by ct0/ct1)
        // rint(t) with t = I0 / Z: (-t + 1.5 * 2^23) is an exact integer shift for |t| < 2^22; lanes beyond are never `go`
        const f32x2 nixm = fma2(I0, s, bc2(12582912.0f));
        const f32x2 ix = sub2(bc2(12582912.0f), nixm);  // = rint(tx), exact
        const f32x2 fx = fma2(I0, s, ix);       // rint(tx) - tx
        const f32x2 niym = fma2(I1, s, bc2(12582912.0f));
        const f32x2 iy = sub2(bc2(12582912.0f), niym);
        const f32x2 fy = fma2(I1, s, iy);
        const f32x2 thr = fma2(ct1_2, s, hct0_2);  // 1/2 - ct0 - ct1 / Z
        float fxl, fxh, fyl, fyh, thl, thh;
        upk2(fx, fxl, fxh), upk2(fy, fyl, fyh), upk2(thr, thl, thh);
        // the reference rounds to the same depth pixel, and the point is certainly inside the image in front of the camera
        const bool gol = (fmaxf(fabsf(fxl), fabsf(fyl)) <= thl) && (ml >= B2q);
        const bool goh = (fmaxf(fabsf(fxh), fabsf(fyh)) <= thh) && (mh >= B2q);
        // gather index: di + 2^23 as an exact float, i.e. the bit pattern 0x4B000000 + di (dm is pre-biased by that constant)
        const f32x2 ixb = add2(ix, bc2(8388608.0f));
        float ixl, ixh, iyl, iyh;
        upk2(ixb, ixl, ixh), upk2(iy, iyl, iyh);
        const f32x2 dbits = fma2(pk2(fminf(iyl, q.yclip), fminf(iyh, q.yclip)), bc2(q.fdw), pk2(fminf(ixl, q.xclip), fminf(ixh, q.xclip)));
        float dbl, dbh;
        upk2(dbits, dbl, dbh);
        const uint32_t bitl = gol ? __float_as_uint(dbl) : 0x4B000000u, bith = goh ? __float_as_uint(dbh) : 0x4B000000u;
        bz[k] = pk2(gol ? Zl : ((ml >= nB) ? qnan : finf), goh ? Zh : ((mh >= nB) ? qnan : finf));
        brl[k] = ldg_u16_nc(reinterpret_cast<const uint16_t*>(dm + 2ull * bitl));
        brh[k] = ldg_u16_nc(reinterpret_cast<const uint16_t*>(dm + 2ull * bith));
      }
      // the frame's gathers are in flight: look the next candidate frame up now
      fl_next = __ffs(fm) - 1;
      g_next = __shfl_sync(0xffffffffu, gmr[r], fl_next & 31);
      fm &= fm - 1u;
      uint32_t wl = 0u, wh = 0u;
#pragma unroll
      for (int k = 0; k < kC9SgPerWarp; ++k) {
        if (!((g >> k) & 1u)) continue;
        // depth: d^ = fl(float(raw) * 0.001f); visible iff z < d is certain when Z^ < d^ - (B_Z + 6u d^)
        const f32x2 d = mul2(add2(pk2(__uint_as_float(0x4B000000u | brl[k]), __uint_as_float(0x4B000000u | brh[k])), bc2(-8388608.0f)), bc2(q.fscale));
        const f32x2 mz = fma2(bc2(3.6e-7f), d, bc2(sc4.z));
        const f32x2 t = sub2(bz[k], d);
        float tl, th, mzl, mzh;
        upk2(t, tl, th), upk2(mz, mzl, mzh);
        const bool visl = tl < -mzl, vish = th < -mzh;
        const bool uncl = !(fabsf(tl) > mzl), unch = !(fabsf(th) > mzh);
        const uint32_t bvl = __ballot_sync(0xffffffffu, visl), bvh = __ballot_sync(0xffffffffu, vish);
        const uint32_t bu = __ballot_sync(0xffffffffu, uncl | unch);
        if (lane == k) wl = bvl, wh = bvh;
        if (bu != 0u) {  // rare, warp-uniform: queue the uncertain lanes (lo half first); worked off once the frame's words are stored
          const uint32_t bul = __ballot_sync(0xffffffffu, uncl), buh = __ballot_sync(0xffffffffu, unch);
          const uint32_t code = ((uint32_t)fi << 8) | ((uint32_t)k << 6) | (uint32_t)lane;
          if (uncl) s_dq[warp][dq + __popc(bul & lt)] = (uint16_t)code;
          dq += __popc(bul);
          if (unch) s_dq[warp][dq + __popc(buh & lt)] = (uint16_t)(code | 32u);
          dq += __popc(buh);
        }
      }
      store_frame(fi, wl, wh);
    }
  }
  if (dq != 0u) flush(1u);
}

// ================================================================================================================
// K2: z-buffer splat in the same structure (msp_zbuffer_render with standard intrinsics).  NOT in the reference (our definition:
// per depth pixel the minimum float32(z_ref) over the points that land on it, pixel by the reference's index rule IH:359-366).
// Classification and the packed float32 projection are K1's; a lane whose depth pixel is certain is QUEUED with that pixel
// (round 1 recomputed z in float64 inline, i.e. at the lane efficiency of "how many of these 32 points are in view": ~15 % on
// random clouds), uncertain lanes are queued without one, and the queue is worked off 32 entries at a time with all lanes busy:
// original coordinates from global memory, z = row 2 of w2c . [x y z 1] in the reference's float64 operation order (certain
// pixel) or the whole exact projection (uncertain), then one atomicMin on the ordered float bits.
// ================================================================================================================
constexpr int kZ9Frames = 32, kZ9Rounds = 1;  // frames per CTA of k_zbuf_cull9: it packs the frame index of a queued point into 5 bits
constexpr uint32_t kZ9NoPixel = 0x7ffffu;  // 19-bit pixel field: "uncertain, run the exact projection" (needs depth_w * depth_h < 2^19 - 1)

template <int DT>
__global__ void __launch_bounds__(kC9Threads, MSP_C9_MINBLOCKS) k_zbuf_cull9(const Cull9Params q) {
  extern __shared__ __align__(128) unsigned char s_raw[];  // CullTilePrep without the instance part
  __shared__ __align__(16) C9Frame s_c[kZ9Frames];
  __shared__ __align__(16) double s_wz[kZ9Frames][4];       // row 2 of every frame's w2c (the z form, float64)
  __shared__ uint32_t s_zq[kC9Warps][96];                   // pixel << 13 | frame in chunk << 8 | k << 6 | half << 5 | lane
  __shared__ __align__(8) uint64_t s_bar;
  const CullParams& p = q.p;
  const CullTilePrep& T = *reinterpret_cast<const CullTilePrep*>(s_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kC9TilePoints;
  const int cnt = (int)min((int64_t)kC9TilePoints, p.n - base);
  const int f0 = blockIdx.y * kZ9Frames;
  const int nf = min(kZ9Frames, p.F - f0);
  const CullTilePrep* gprep = q.prep + blockIdx.x;

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    constexpr uint32_t bytes = (uint32_t)offsetof(CullTilePrep, pt_slot);
    mbar_arrive_expect_tx(&s_bar, bytes);
    bulk_g2s(s_raw, gprep, bytes, &s_bar);
  }
  if (tid < nf) {
    const float centre[3] = {__ldg(&gprep->centre[0]), __ldg(&gprep->centre[1]), __ldg(&gprep->centre[2])};
    const float extf[3] = {__ldg(&gprep->ext[0]), __ldg(&gprep->ext[1]), __ldg(&gprep->ext[2])};
    const double* w = p.w2c + (int64_t)(f0 + tid) * 16;
    c9_frame_consts(w, p.K, p.g, centre, extf, nullptr, s_c[tid]);
    s_wz[tid][0] = w[8], s_wz[tid][1] = w[9], s_wz[tid][2] = w[10], s_wz[tid][3] = w[11];
  }
  mbar_wait(&s_bar, 0);
  __syncthreads();

  const uint32_t lt = (1u << lane) - 1u;
  uint32_t s_lane = smem_u32(s_raw) + (uint32_t)MSP_C9_SG(warp, 0) * 768u + (uint32_t)lane * 8u;
  asm volatile("" : "+r"(s_lane));
  constexpr uint32_t kSgStride = (uint32_t)(MSP_C9_SG(0, 1) - MSP_C9_SG(0, 0)) * 768u;

  // ---- A: classification, lane = frame (as in K1)
  uint32_t gmr[kZ9Rounds];
#pragma unroll
  for (int r = 0; r < kZ9Rounds; ++r) {
    const int fi = 32 * r + lane;
    uint32_t gm = 0u;
    if (fi < nf) {
      const C9Frame& c = s_c[fi];
      const float4 a0 = *reinterpret_cast<const float4*>(c.a0), a2 = *reinterpret_cast<const float4*>(c.a2);
      const float4 q1 = *reinterpret_cast<const float4*>(c.q1), q3 = *reinterpret_cast<const float4*>(c.q3);
      const float nB = -c.B;
#pragma unroll
      for (int k = 0; k < kC9SgPerWarp; ++k) {
        const float4 bc = T.box[MSP_C9_SG(warp, k)][0], bh = T.box[MSP_C9_SG(warp, k)][1];
        const float v0 = fmaf(a0.x, bc.x, fmaf(a0.y, bc.y, fmaf(a0.z, bc.z, a0.w)));
        const float r0 = __fmaf_ru(fabsf(a0.x), bh.x, __fmaf_ru(fabsf(a0.y), bh.y, __fmul_ru(fabsf(a0.z), bh.z)));
        const float v1 = fmaf(q1.x, bc.x, fmaf(q1.y, bc.y, fmaf(q1.z, bc.z, q1.w)));
        const float r1 = __fmaf_ru(fabsf(q1.x), bh.x, __fmaf_ru(fabsf(q1.y), bh.y, __fmul_ru(fabsf(q1.z), bh.z)));
        const float v2 = fmaf(a2.x, bc.x, fmaf(a2.y, bc.y, fmaf(a2.z, bc.z, a2.w)));
        const float r2 = __fmaf_ru(fabsf(a2.x), bh.x, __fmaf_ru(fabsf(a2.y), bh.y, __fmul_ru(fabsf(a2.z), bh.z)));
        const float v3 = fmaf(q3.x, bc.x, fmaf(q3.y, bc.y, fmaf(q3.z, bc.z, q3.w)));
        const float r3 = __fmaf_ru(fabsf(q3.x), bh.x, __fmaf_ru(fabsf(q3.y), bh.y, __fmul_ru(fabsf(q3.z), bh.z)));
        const bool cand = (__fadd_ru(v0, r0) >= nB) && (__fadd_ru(v1, r1) >= nB) && (__fadd_ru(v2, r2) >= nB) && (__fadd_ru(v3, r3) >= nB);
        if (cand) gm |= 1u << k;
      }
    }
    gmr[r] = gm;
  }

  uint32_t zq = 0;  // warp-uniform fill of s_zq[warp]
  auto splat = [&](uint32_t e) {  // one queued point on this lane
    const uint32_t pix = e >> 13;
    const int fi = (int)((e >> 8) & 31u), k = (int)((e >> 6) & 3u), sl = (int)(e & 63u);
    const int idx = MSP_C9_SG(warp, k) * 64 + sl;
    if (idx >= cnt) return;
    double x, y, z;
    PtTraits<DT>::load(p.pts, base + idx, x, y, z);
    unsigned int* zrow = reinterpret_cast<unsigned int*>(p.zbuf + (int64_t)(f0 + fi) * p.depth_stride);
    if (pix != kZ9NoPixel) {
      // the pixel is certain; the stored value must be float32(z_ref): z alone, in the reference's operation order (IH:60,63)
      const double zref = row_dot_h(s_wz[fi], x, y, z);
      atomicMin(zrow + pix, __float_as_uint(__double2float_rn(zref)));
    } else {
      const double* w = p.w2c + (int64_t)(f0 + fi) * 16;
      const Proj pr = (s_c[fi].std_frame != 0) ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
      if (in_image(pr, p.g) && pr.z > 0.0) atomicMin(zrow + depth_index(pr, p.g), __float_as_uint(__double2float_rn(pr.z)));
    }
  };
  auto flush = [&](uint32_t keep_below) {
    __syncwarp();
    while (zq >= keep_below && zq != 0u) {
      const uint32_t nb = zq < 32u ? zq : 32u;
      if (lane < nb) splat(s_zq[warp][zq - nb + lane]);
      zq -= nb;
    }
    __syncwarp();
  };

#pragma unroll 1
  for (int r = 0; r < kZ9Rounds; ++r) {
    uint32_t fm = __ballot_sync(0xffffffffu, gmr[r] != 0u);
#pragma unroll 1
    while (fm != 0u) {
      const int fl = __ffs(fm) - 1;
      fm &= fm - 1u;
      uint32_t g = __shfl_sync(0xffffffffu, gmr[r], fl);
      const int fi = 32 * r + fl;
      const uint32_t s_cf = smem_u32(&s_c[fi]);
      const float4 a0 = lds_f4(s_cf + 0), a2 = lds_f4(s_cf + 16), az = lds_f4(s_cf + 32), sc4 = lds_f4(s_cf + 80);
      const f32x2 ct1_2 = bc2(sc4.x), hct0_2 = bc2(sc4.y);
      const float B2q = sc4.w + sc4.w, nB = -sc4.w;
#pragma unroll 1
      while (g != 0u) {
        const int k = __ffs(g) - 1;
        g &= g - 1u;
        const uint32_t pa = s_lane + kSgStride * (uint32_t)k;
        const f32x2 xx = lds64(pa), yy = lds64(pa + 256u), zz = lds64(pa + 512u);
        const f32x2 I0 = fma2(bc2(a0.x), xx, fma2(bc2(a0.y), yy, fma2(bc2(a0.z), zz, bc2(a0.w))));
        const f32x2 I1 = fma2(bc2(a2.x), xx, fma2(bc2(a2.y), yy, fma2(bc2(a2.z), zz, bc2(a2.w))));
        const f32x2 Zp = fma2(bc2(az.x), xx, fma2(bc2(az.y), yy, fma2(bc2(az.z), zz, bc2(az.w))));
        const f32x2 nQ1 = fma2(bc2(q.nfdw), Zp, I0);
        const f32x2 nQ3 = fma2(bc2(q.nfdh), Zp, I1);
        float I0l, I0h, I1l, I1h, Zl, Zh, nq1l, nq1h, nq3l, nq3h;
        upk2(I0, I0l, I0h), upk2(I1, I1l, I1h), upk2(Zp, Zl, Zh), upk2(nQ1, nq1l, nq1h), upk2(nQ3, nq3l, nq3h);
        const float ml = fminf(fminf(I0l, I1l), -fmaxf(nq1l, nq3l));
        const float mh = fminf(fminf(I0h, I1h), -fmaxf(nq1h, nq3h));
        const f32x2 s = pk2(rcp_approx(-Zl), rcp_approx(-Zh));
        const f32x2 nixm = fma2(I0, s, bc2(12582912.0f));
        const f32x2 ix = sub2(bc2(12582912.0f), nixm);
        const f32x2 fx = fma2(I0, s, ix);
        const f32x2 niym = fma2(I1, s, bc2(12582912.0f));
        const f32x2 iy = sub2(bc2(12582912.0f), niym);
        const f32x2 fy = fma2(I1, s, iy);
        const f32x2 thr = fma2(ct1_2, s, hct0_2);
        float fxl, fxh, fyl, fyh, thl, thh;
        upk2(fx, fxl, fxh), upk2(fy, fyl, fyh), upk2(thr, thl, thh);
        const bool gol = (fmaxf(fabsf(fxl), fabsf(fyl)) <= thl) && (ml >= B2q);
        const bool goh = (fmaxf(fabsf(fxh), fabsf(fyh)) <= thh) && (mh >= B2q);
        const bool uncl = !gol && (ml >= nB), unch = !goh && (mh >= nB);
        const f32x2 ixb = add2(ix, bc2(8388608.0f));
        float ixl, ixh, iyl, iyh;
        upk2(ixb, ixl, ixh), upk2(iy, iyl, iyh);
        const f32x2 dbits = fma2(pk2(fminf(iyl, q.yclip), fminf(iyh, q.yclip)), bc2(q.fdw), pk2(fminf(ixl, q.xclip), fminf(ixh, q.xclip)));
        float dbl, dbh;
        upk2(dbits, dbl, dbh);
        // queue every lane that may land on a pixel: certain ones with the pixel, uncertain ones with the "no pixel" mark
        const uint32_t bql = __ballot_sync(0xffffffffu, gol | uncl), bqh = __ballot_sync(0xffffffffu, goh | unch);
        const uint32_t code = ((uint32_t)fi << 8) | ((uint32_t)k << 6) | (uint32_t)lane;
        const uint32_t el = ((gol ? (__float_as_uint(dbl) - 0x4B000000u) : kZ9NoPixel) << 13) | code;
        const uint32_t eh = ((goh ? (__float_as_uint(dbh) - 0x4B000000u) : kZ9NoPixel) << 13) | code | 32u;
        // (splatting a dense half in place instead of queueing it was measured 3.5x slower on mesh-ordered clouds: divergent
        //  float64 code inside the projection loop; profiles/r2/k1_tuning.md)
        if (gol | uncl) s_zq[warp][zq + __popc(bql & lt)] = el;
        zq += __popc(bql);
        if (goh | unch) s_zq[warp][zq + __popc(bqh & lt)] = eh;
        zq += __popc(bqh);
        if (zq >= 32u) flush(32u);
      }
    }
  }
  if (zq != 0u) flush(1u);
}

}  // namespace msp
