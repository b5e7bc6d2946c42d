// libmsp.so — hand-written sm_100a kernels + C ABI for the MultiSPA geometry hot path.
// Entry points and the reference lines they replace are documented in include/msp.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "msp.h"
#include "msp_device.cuh"
#include "msp_cull.cuh"
#include "msp_cull9.cuh"
#include "msp_tables.cuh"

using namespace msp;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return MSP_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return MSP_ELAUNCH;
}
#define REQUIRE(cond, msg)                                    \
  do {                                                        \
    if (!(cond)) return fail(MSP_EINVAL, msg);          \
  } while (0)

static unsigned long long* g_cull_stats = nullptr;  // device pointer set by msp_debug_set_cull_stats (profiling aid)
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// K0: pose prologue (one thread per frame)
// ------------------------------------------------------------------------------------------------
__global__ void k_pose_prepare(const double* __restrict__ c2w, const double* __restrict__ axis_align, int F,
                               double* __restrict__ E_out, double* __restrict__ w2c_out,
                               double* __restrict__ yawpitch_out, int* __restrict__ valid_out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double C[16], E[16], inv[16];
  bool finite = true;
  for (int i = 0; i < 16; ++i) {
    C[i] = c2w[(int64_t)f * 16 + i];
    finite &= isfinite(C[i]);  // info_handler.py:414-416
  }
  if (axis_align != nullptr) {
    double A[16];
    for (int i = 0; i < 16; ++i) A[i] = axis_align[i];
    matmul4(A, C, E);  // info_handler.py:122
  } else {
    for (int i = 0; i < 16; ++i) E[i] = C[i];
  }
  bool ok = finite && inverse4(E, inv);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = 0; i < 16; ++i) {
    E_out[(int64_t)f * 16 + i] = E[i];
    w2c_out[(int64_t)f * 16 + i] = ok ? inv[i] : qnan;
  }
  // calculate_frames_relations.py:86-100 — viewing direction is the third column of the rotation block
  const double zx = E[2], zy = E[6], zz = E[10];
  const double kDeg = 180.0 / 3.14159265358979323846;
  const double nrm = sqrt(__fma_rn(zz, zz, __fma_rn(zy, zy, __dmul_rn(zx, zx))));
  yawpitch_out[(int64_t)f * 2 + 0] = atan2(zy, zx) * kDeg;
  yawpitch_out[(int64_t)f * 2 + 1] = asin(zz / nrm) * kDeg;
  valid_out[f] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// K1 / K2: per-(frame, point) projection with either the depth-map visibility test (bit mask out)
// or a z-buffer splat.  A CTA stages one tile of points in shared memory with a bulk async copy,
// keeps each lane's points in registers as float64 and walks a chunk of frames over them.
// ------------------------------------------------------------------------------------------------
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWordsPerWarp = 4;                          // 32-bit mask words produced per warp per frame
constexpr int kTilePoints = kWarps * kWordsPerWarp * 32;  // 1024
constexpr int kFramesPerCta = 32;

struct FrameLoopParams {
  const void* pts;
  int64_t n;
  const double* w2c;  // (F,16)
  int32_t F;
  const uint16_t* depth;  // (F, dh*dw)        [visibility]
  float* zbuf;            // (F, dh*dw)        [z-buffer]
  int64_t depth_stride;   // dh*dw
  uint32_t* mask;
  int64_t mask_stride;
  double* xy;
  double* z;
  Intrinsics K;
  ImageGeom g;
  const uint16_t* inst;  // fused object counting (fast path only)
  uint32_t* counts;
  int32_t n_objects;
};

enum { MODE_VIS = 0, MODE_ZBUF = 1 };

template <int DT, bool KSTD, int MODE, bool MAT>
__global__ void __launch_bounds__(kThreads) k_frame_loop(const FrameLoopParams p) {
  __shared__ __align__(128) unsigned char s_pts[kTilePoints * PtTraits<DT>::kBytes];
  __shared__ __align__(16) double s_w[kFramesPerCta * 16];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kTilePoints;
  const int cnt = (int)min((int64_t)kTilePoints, p.n - base);
  const int f0 = blockIdx.y * kFramesPerCta;
  const int nf = min(kFramesPerCta, p.F - f0);

  // ---- stage the point tile: one elected thread issues the bulk copy, everybody waits on the mbarrier
  const uint32_t bytes = (uint32_t)cnt * PtTraits<DT>::kBytes;
  const uint32_t bulk = bytes & ~15u;
  const unsigned char* gsrc = static_cast<const unsigned char*>(p.pts) + base * PtTraits<DT>::kBytes;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s_bar, bulk);
    if (bulk) bulk_g2s(s_pts, gsrc, bulk, &s_bar);
  }
  if (tid < (int)(bytes - bulk)) s_pts[bulk + tid] = gsrc[bulk + tid];
  for (int i = tid; i < nf * 16; i += kThreads) s_w[i] = p.w2c[(int64_t)f0 * 16 + i];
  mbar_wait(&s_bar, 0);
  __syncthreads();

  double px[kWordsPerWarp], py[kWordsPerWarp], pz[kWordsPerWarp];
  bool ok[kWordsPerWarp];
#pragma unroll
  for (int j = 0; j < kWordsPerWarp; ++j) {
    const int idx = warp * (kWordsPerWarp * 32) + j * 32 + lane;
    ok[j] = idx < cnt;
    px[j] = py[j] = pz[j] = 0.0;
    if (ok[j]) PtTraits<DT>::load(s_pts, idx, px[j], py[j], pz[j]);
  }
  const int64_t word0 = (int64_t)blockIdx.x * (kTilePoints / 32) + warp * kWordsPerWarp;

  for (int fi = 0; fi < nf; ++fi) {
    const double* w = s_w + fi * 16;
    const int64_t f = f0 + fi;
    const bool fstd = KSTD && (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
    uint32_t myword = 0;
#pragma unroll
    for (int j = 0; j < kWordsPerWarp; ++j) {
      Proj pr = fstd ? project_exact<true>(w, p.K, px[j], py[j], pz[j])
                     : project_exact<false>(w, p.K, px[j], py[j], pz[j]);
      const bool cand = ok[j] && in_image(pr, p.g) && (pr.z > 0.0);
      if (MODE == MODE_VIS) {
        bool vis = false;
        if (cand) vis = depth_test(pr.z, ldg_u16_nc(p.depth + f * p.depth_stride + depth_index(pr, p.g)), p.g);
        const uint32_t b = __ballot_sync(0xffffffffu, vis);
        if (lane == j) myword = b;
      } else {
        if (cand)
          atomicMin(reinterpret_cast<unsigned int*>(p.zbuf + f * p.depth_stride) + depth_index(pr, p.g),
                    __float_as_uint(__double2float_rn(pr.z)));
      }
      if (MAT && ok[j]) {
        const int64_t gi = f * p.n + base + warp * (kWordsPerWarp * 32) + j * 32 + lane;
        reinterpret_cast<double2*>(p.xy)[gi] = make_double2(pr.u, pr.v);
        p.z[gi] = pr.z;
      }
    }
    if (MODE == MODE_VIS) {
      if (lane < kWordsPerWarp && word0 + lane < p.mask_stride) p.mask[f * p.mask_stride + word0 + lane] = myword;
    }
  }
}

__global__ void k_fill_u32(uint32_t* __restrict__ dst, int64_t count, uint32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = i; k < count; k += stride) dst[k] = v;
}

// Synthetic sensor depth from a z-buffer: raw = min(floor(z / scale) + 1 + margin, 65535) as uint16, 0 where no point landed
// (z = +inf).  A pure streaming pass (4 B in, 2 B out per pixel), so it is written for bandwidth: 8 pixels per thread with two
// 128-bit loads and one 128-bit store, and the float64 division of the definition is only executed inside a guard band:
// qf = fl32(z * fl32(1/scale)) is within 1.2e-7 |q| of the exact quotient, so whenever its fractional part is farther than
// 4e-7 |q| + 1e-6 from both neighbouring integers, floor(qf) IS floor(fl64(z / scale)) — otherwise the lane evaluates the
// definition itself (__ddiv_rn + floor).  Bit-identical to the all-float64 formula (tests/test_gpu_parity.py), ~3x faster.
__device__ __forceinline__ uint32_t depth_raw_u16(float zf, float rs, double depth_scale, float add) {
  if (!(fabsf(zf) <= 3.4028234663852886e38f)) return 0u;  // +-inf / NaN: no point landed on this pixel
  const float qf = zf * rs;
  const float fl = floorf(qf);
  const float frac = qf - fl;  // exact
  const float delta = fmaf(fabsf(qf), 4e-7f, 1e-6f);
  float mm;
  if (frac > delta && frac < 1.0f - delta && fabsf(qf) < 4194304.0f) {
    mm = fl + add;  // integers below 2^23: exact
  } else {
    mm = (float)fmin(floor(__ddiv_rn((double)zf, depth_scale)) + (double)add, 65535.0);  // <= 65535 is exact in float32
  }
  mm = fminf(mm, 65535.0f);
  return (uint32_t)(uint16_t)(int)fmaxf(mm, 0.0f);  // a negative z never comes out of the z-buffer (only z > 0 is splatted)
}

__global__ void __launch_bounds__(256) k_depth_to_u16(const float* __restrict__ zbuf, int64_t count, int margin_mm,
                                                      double depth_scale, uint16_t* __restrict__ out) {
  const float rs = (float)(1.0 / depth_scale);
  const float add = (float)(1 + margin_mm);
  const int64_t n8 = count >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4* __restrict__ in4 = reinterpret_cast<const float4*>(zbuf);
  uint4* __restrict__ out4 = reinterpret_cast<uint4*>(out);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n8; k += stride) {
    const float4 a = __ldcs(in4 + 2 * k), b = __ldcs(in4 + 2 * k + 1);  // streamed once: evict-first
    uint4 r;
    r.x = depth_raw_u16(a.x, rs, depth_scale, add) | (depth_raw_u16(a.y, rs, depth_scale, add) << 16);
    r.y = depth_raw_u16(a.z, rs, depth_scale, add) | (depth_raw_u16(a.w, rs, depth_scale, add) << 16);
    r.z = depth_raw_u16(b.x, rs, depth_scale, add) | (depth_raw_u16(b.y, rs, depth_scale, add) << 16);
    r.w = depth_raw_u16(b.z, rs, depth_scale, add) | (depth_raw_u16(b.w, rs, depth_scale, add) << 16);
    out4[k] = r;
  }
  // tail (count not a multiple of 8) and nothing else: one thread per element
  for (int64_t k = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    out[k] = (uint16_t)depth_raw_u16(zbuf[k], rs, depth_scale, add);
}

// ------------------------------------------------------------------------------------------------
// K3/K4: the per-pair stage.  One CTA per pair; each warp owns a contiguous quarter of the mask row.
// ------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 128;
constexpr int kPairWarps = kPairThreads / 32;

struct PairParams {
  const uint32_t* mask;
  int64_t n;
  int64_t mask_stride;
  const int32_t* pairs;
  int64_t P;
  const uint32_t* sel;
  int32_t sel_mode;
  const void* pts;
  const double* E;
  const double* w2c;
  const double* yawpitch;
  const uint16_t* depth;
  int64_t depth_stride;
  Intrinsics K;
  ImageGeom g;
  msp_pair_record* rec;
  uint32_t* inter_out;
  uint32_t* uni_out;
  double* overlap_out;
};

__device__ __forceinline__ uint32_t popc4(const uint4 v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
__device__ __forceinline__ uint4 and4(const uint4 a, const uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 or4(const uint4 a, const uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }

template <int DT, bool KSTD, bool ANNOTATE>
__global__ void __launch_bounds__(kPairThreads) k_pair(const PairParams p) {
  __shared__ uint32_t s_inter[kPairWarps], s_uni[kPairWarps];
  __shared__ int32_t s_point;
  __shared__ __align__(16) msp_pair_record s_rec;

  const int64_t pid = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t fa = p.pairs[2 * pid], fb = p.pairs[2 * pid + 1];
  const uint4* __restrict__ A = reinterpret_cast<const uint4*>(p.mask + (int64_t)fa * p.mask_stride);
  const uint4* __restrict__ B = reinterpret_cast<const uint4*>(p.mask + (int64_t)fb * p.mask_stride);
  const int nq = (int)(((p.n + 31) / 32 + 3) / 4);       // uint4 quads holding valid words
  const int qpw = (nq + kPairWarps - 1) / kPairWarps;    // quads per warp segment
  const int lo = warp * qpw, hi = min(nq, lo + qpw);

  uint32_t ci = 0, cu = 0;
  for (int q = lo + lane; q < hi; q += 32) {
    const uint4 a = __ldg(A + q), b = __ldg(B + q);
    ci += popc4(and4(a, b));
    cu += popc4(or4(a, b));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ci += __shfl_xor_sync(0xffffffffu, ci, o);
    cu += __shfl_xor_sync(0xffffffffu, cu, o);
  }
  if (lane == 0) s_inter[warp] = ci, s_uni[warp] = cu;
  if (tid == 0) s_point = -1;
  __syncthreads();
  uint32_t inter = 0, uni = 0;
#pragma unroll
  for (int w = 0; w < kPairWarps; ++w) inter += s_inter[w], uni += s_uni[w];

  if (!ANNOTATE) {
    if (tid == 0) {
      p.inter_out[pid] = inter;
      p.uni_out[pid] = uni;
      p.overlap_out[pid] = __dmul_rn(__ddiv_rn((double)inter, (double)uni), 100.0);  // CFR:136
    }
    return;
  }

  // ---- k-th smallest common point id (np.intersect1d order = ascending word, ascending bit)
  const uint32_t sv = p.sel[pid];
  const uint64_t k64 = (p.sel_mode == 1) ? (((uint64_t)sv * (uint64_t)inter) >> 32) : (uint64_t)sv;
  if (k64 < inter) {
    uint32_t k = (uint32_t)k64, before = 0;
    int target = 0;
    for (int w = 0; w < kPairWarps; ++w) {
      if (k >= before + s_inter[w]) before += s_inter[w], target = w + 1;
      else break;
    }
    if (warp == target) {
      k -= before;
      for (int q0 = lo; q0 < hi; q0 += 32) {  // warp-uniform loop
        const int q = q0 + lane;
        uint4 c = make_uint4(0, 0, 0, 0);
        if (q < hi) c = and4(__ldg(A + q), __ldg(B + q));
        const uint32_t cnt = popc4(c);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (k < total) {
          const uint32_t excl = incl - cnt;
          if (k >= excl && k < incl) {
            uint32_t r = k - excl;
            const uint32_t wv[4] = {c.x, c.y, c.z, c.w};
            int wi = 0;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const uint32_t pc = __popc(wv[wi]);
              if (r >= pc && wi == t) r -= pc, wi = t + 1;
            }
            const uint32_t word = wi == 0 ? c.x : wi == 1 ? c.y : wi == 2 ? c.z : c.w;
            const int bit = __fns(word, 0, (int)r + 1);
            s_point = (q * 4 + wi) * 32 + bit;
          }
          break;
        }
        k -= total;
      }
    }
  }
  __syncthreads();
  const int32_t point = s_point;

  if (warp == 0) {
    if (lane < 2) {  // annotate the point in frame 1 / frame 2
      const int32_t f = lane == 0 ? fa : fb;
      int vis = 0, pmx = 0, pmy = 0, dmm = 0, ux = 0, vy = 0;
      if (point >= 0) {
        double x, y, z;
        PtTraits<DT>::load(p.pts, point, x, y, z);
        const double* w = p.w2c + (int64_t)f * 16;
        const bool fstd = KSTD && (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
        Proj pr;
        const uint16_t* dm = p.depth + (int64_t)f * p.depth_stride;
        // per-point call of the reference (IH:291-305): one-column matmul order
        (void)fstd;
        vis = visible_exact<false, true>(w, p.K, p.g, dm, x, y, z, pr);
        pmx = permille(pr.u, p.g.img_w);
        pmy = permille(pr.v, p.g.img_h);
        dmm = depth_mm(pr.z);
        ux = __double2int_rz(pr.u);
        vy = __double2int_rz(pr.v);
      }
      s_rec.permille[2 * lane] = pmx;
      s_rec.permille[2 * lane + 1] = pmy;
      s_rec.depth_mm[lane] = dmm;
      s_rec.pix[2 * lane] = ux;
      s_rec.pix[2 * lane + 1] = vy;
      const int other = __shfl_xor_sync(0x3u, vis, 1);
      if (lane == 0) s_rec.flags = (vis ? 1 : 0) | (other ? 2 : 0);
    } else if (lane == 2) {  // pose relations
      const double* E1 = p.E + (int64_t)fa * 16;
      const double* E2 = p.E + (int64_t)fb * 16;
      const double dx = E2[3] - E1[3], dy = E2[7] - E1[7], dz = E2[11] - E1[11];
      s_rec.inter = inter;
      s_rec.uni = uni;
      s_rec.overlap = __dmul_rn(__ddiv_rn((double)inter, (double)uni), 100.0);
      s_rec.distance = sqrt(__fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx))));  // CFR:183
      s_rec.dyaw = p.yawpitch[2 * (int64_t)fb] - p.yawpitch[2 * (int64_t)fa];        // CFR:181
      s_rec.dpitch = p.yawpitch[2 * (int64_t)fb + 1] - p.yawpitch[2 * (int64_t)fa + 1];
      const double* w1 = p.w2c + (int64_t)fa * 16;  // inv(E1) @ E2, translation column (CME:185-189)
#pragma unroll
      for (int i = 0; i < 3; ++i) s_rec.rel_disp[i] = row_dot4(w1 + 4 * i, E2[3], E2[7], E2[11], E2[15]);
      s_rec.point_id = point;
    } else if (lane == 3) {
      s_rec.reserved[0] = s_rec.reserved[1] = s_rec.reserved[2] = s_rec.reserved[3] = 0;
    }
    __syncwarp();
    reinterpret_cast<uint32_t*>(p.rec + pid)[lane] = reinterpret_cast<const uint32_t*>(&s_rec)[lane];
  }
}

// ------------------------------------------------------------------------------------------------
// point annotation requests (one thread each)
// ------------------------------------------------------------------------------------------------
struct AnnotParams {
  const void* pts;
  int64_t n;
  const int32_t* point_ids;
  const int32_t* frame_ids;
  int64_t M;
  const double* w2c;
  const uint16_t* depth;
  int64_t depth_stride;
  Intrinsics K;
  ImageGeom g;
  double* uvz;
  int32_t* permille;
  int32_t* depth_mm;
  int32_t* pix;
  uint8_t* visible;
  int batch_order;  // 1: N >= 2 gemm order (batch projection of one frame), 0: single-point order (IH:291-305)
};

template <int DT, bool KSTD>
__global__ void k_annotate(const AnnotParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M) return;
  const int64_t pt = p.point_ids ? (int64_t)p.point_ids[i] : i;
  const int32_t f = p.frame_ids ? p.frame_ids[i] : 0;
  double x, y, z;
  PtTraits<DT>::load(p.pts, pt, x, y, z);
  const double* w = p.w2c + (int64_t)f * 16;
  const bool fstd = KSTD && (w[12] == 0.0) && (w[13] == 0.0) && (w[14] == 0.0) && (w[15] == 1.0);
  Proj pr;
  bool vis;
  if (p.batch_order) {
    pr = fstd ? project_exact<true>(w, p.K, x, y, z) : project_exact<false>(w, p.K, x, y, z);
    vis = in_image(pr, p.g) && pr.z > 0.0;
    if (vis && p.depth != nullptr)
      vis = depth_test(pr.z, ldg_u16_nc(p.depth + (int64_t)f * p.depth_stride + depth_index(pr, p.g)), p.g);
  } else if (p.depth != nullptr) {
    vis = visible_exact<false, true>(w, p.K, p.g, p.depth + (int64_t)f * p.depth_stride, x, y, z, pr);
  } else {
    pr = project_exact<false, true>(w, p.K, x, y, z);
    vis = in_image(pr, p.g) && pr.z > 0.0;
  }
  if (p.uvz) p.uvz[3 * i] = pr.u, p.uvz[3 * i + 1] = pr.v, p.uvz[3 * i + 2] = pr.z;
  if (p.permille) p.permille[2 * i] = permille(pr.u, p.g.img_w), p.permille[2 * i + 1] = permille(pr.v, p.g.img_h);
  if (p.depth_mm) p.depth_mm[i] = depth_mm(pr.z);
  if (p.pix) p.pix[2 * i] = __double2int_rz(pr.u), p.pix[2 * i + 1] = __double2int_rz(pr.v);
  if (p.visible) p.visible[i] = vis ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// the reference's two-step API: visibility of already projected points (IH:337-386), one thread each
// ------------------------------------------------------------------------------------------------
__global__ void k_check_visibility(const double* __restrict__ xy, const double* __restrict__ z, int64_t n,
                                   const uint16_t* __restrict__ depth, const ImageGeom g, int mode,
                                   uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Proj pr;
  pr.u = xy[2 * i], pr.v = xy[2 * i + 1], pr.z = (z != nullptr) ? z[i] : 0.0;
  bool ok = true;
  if (mode & 1) ok = in_image(pr, g);  // IH:342-343
  if (mode & 2) {                      // IH:359-371 — evaluated for every point like the reference (clip, no bounds AND)
    // np.round(...).astype(int) then clip: out-of-range / non-finite values saturate before the clip
    const double tx = rint(__dmul_rn(pr.u, g.sx)), ty = rint(__dmul_rn(pr.v, g.sy));
    // (x86 float64->int64 conversion of NaN, +-inf and |x| >= 2^63 yields INT64_MIN, which the clip turns into 0)
    const double kBig = 9223372036854775808.0;
    const int dx = (tx >= kBig) ? 0 : (tx >= (double)g.depth_w) ? g.depth_w - 1 : (tx > 0.0 ? (int)tx : 0);
    const int dy = (ty >= kBig) ? 0 : (ty >= (double)g.depth_h) ? g.depth_h - 1 : (ty > 0.0 ? (int)ty : 0);
    ok = ok && depth_test(pr.z, __ldg(depth + (int64_t)dy * g.depth_w + dx), g);
  }
  out[i] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// per-(frame, object) visible-point counts: one CTA per frame, shared-memory histogram,
// warp-aggregated by instance id (neighbouring mesh vertices mostly share one)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_object_counts(const uint32_t* __restrict__ mask, int64_t n, int64_t mask_stride,
                                                       const uint16_t* __restrict__ inst, int n_objects,
                                                       uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t s_hist[];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i <= n_objects; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t* row = mask + (int64_t)f * mask_stride;
  const int nwords = (int)((n + 31) / 32);
  // each warp fetches 32 mask words with one coalesced load, then visits only the non-empty ones
  for (int w0 = warp * 32; w0 < nwords; w0 += (blockDim.x / 32) * 32) {
    const int wd = w0 + lane;
    const uint32_t mine = (wd < nwords) ? __ldg(row + wd) : 0u;
    uint32_t nz = __ballot_sync(0xffffffffu, mine != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t bits = __shfl_sync(0xffffffffu, mine, src);
      const int64_t pt = (int64_t)(w0 + src) * 32 + lane;
      const uint32_t id = (pt < n) ? (uint32_t)__ldg(inst + pt) : 0u;
      const bool active = ((bits >> lane) & 1u) && id != 0u && id <= (uint32_t)n_objects;
      const uint32_t m = __ballot_sync(0xffffffffu, active);
      if (active) {
        const uint32_t grp = __match_any_sync(m, id);
        if (lane == __ffs(grp) - 1) atomicAdd(&s_hist[id], (uint32_t)__popc(grp));
      }
    }
  }
  __syncthreads();
  for (int o = tid; o < n_objects; o += blockDim.x) counts[(int64_t)f * n_objects + o] = s_hist[o + 1];
}

// ------------------------------------------------------------------------------------------------
// per-(frame, object) axis-aligned extent of the VISIBLE object points: min/max of x,y,z.
// single_object_coverage_finder.py:57-65 takes max(coords) - min(coords) over unions of per-image masks; the extent
// of a union is (max of the per-image maxima) - (min of the per-image minima), so this table replaces every
// union-mask pass of the coverage search (SOCF:76-220) by scalar look-ups.  One CTA per frame.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long d2ord(double v) {
  const long long b = __double_as_longlong(v);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double ord2d(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }

template <int DT>
__global__ void __launch_bounds__(256) k_object_extents(const uint32_t* __restrict__ mask, int64_t n, int64_t mask_stride,
                                                        const void* __restrict__ pts, const uint16_t* __restrict__ inst,
                                                        int n_objects, double* __restrict__ out) {
  extern __shared__ long long s_ext[];  // [n_objects][6]: min xyz, max xyz as ordered keys
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long kPosInf = d2ord(__longlong_as_double(0x7ff0000000000000LL));
  const long long kNegInf = d2ord(__longlong_as_double(0xfff0000000000000LL));
  for (int i = tid; i < n_objects * 6; i += blockDim.x) s_ext[i] = (i % 6 < 3) ? kPosInf : kNegInf;
  __syncthreads();
  const uint32_t* row = mask + (int64_t)f * mask_stride;
  const int nwords = (int)((n + 31) / 32);
  for (int w0 = warp * 32; w0 < nwords; w0 += (blockDim.x / 32) * 32) {
    const int wd = w0 + lane;
    const uint32_t mine = (wd < nwords) ? __ldg(row + wd) : 0u;
    uint32_t nz = __ballot_sync(0xffffffffu, mine != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t bits = __shfl_sync(0xffffffffu, mine, src);
      const int64_t pt = (int64_t)(w0 + src) * 32 + lane;
      if (((bits >> lane) & 1u) && pt < n) {
        const uint32_t id = __ldg(inst + pt);
        if (id != 0u && id <= (uint32_t)n_objects) {
          double x, y, z;
          PtTraits<DT>::load(pts, pt, x, y, z);
          long long* e = s_ext + (size_t)(id - 1) * 6;
          if (x == x && y == y && z == z) {
            atomicMin(e + 0, d2ord(x)), atomicMin(e + 1, d2ord(y)), atomicMin(e + 2, d2ord(z));
            atomicMax(e + 3, d2ord(x)), atomicMax(e + 4, d2ord(y)), atomicMax(e + 5, d2ord(z));
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n_objects * 6; i += blockDim.x) out[(int64_t)f * n_objects * 6 + i] = ord2d(s_ext[i]);
}

// ------------------------------------------------------------------------------------------------
// TAPVid-3D tracks
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_track_project(const float* __restrict__ tracks, const double* __restrict__ w2c,
                                                       const double* __restrict__ intr, int T, int N, double img_h,
                                                       double img_w, double* __restrict__ world,
                                                       double* __restrict__ uvn, int32_t* __restrict__ pm,
                                                       uint8_t* __restrict__ valid) {
  __shared__ double s_c2w[16];
  const int t = blockIdx.x;
  if (threadIdx.x == 0) {
    double m[16], inv[16];
    for (int i = 0; i < 16; ++i) m[i] = w2c[(int64_t)t * 16 + i];
    const bool ok = inverse4(m, inv);
    for (int i = 0; i < 16; ++i) s_c2w[i] = ok ? inv[i] : __longlong_as_double(0x7ff8000000000000LL);
  }
  __syncthreads();
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int64_t i = (int64_t)t * N + n;
    const double x = tracks[3 * i], y = tracks[3 * i + 1], z = tracks[3 * i + 2];
    if (world) {  // single_object_movement_engine_coord.py:446-451
#pragma unroll
      // np.einsum('nij,nkj->nki') accumulates these 4-term products as (a0b0 + a2b2) + (a1b1 + a3b3), no FMA
      for (int r = 0; r < 3; ++r) world[3 * i + r] = row_dot_h_single(s_c2w + 4 * r, x, y, z);
    }
    // single_object_movement_engine_coord.py:293-315 (float64 throughout, as under the pinned numpy 1.26)
    const double zg = __dadd_rn(z, 1e-8);
    const double u = __dadd_rn(__ddiv_rn(__dmul_rn(fx, x), zg), cx);
    const double v = __dadd_rn(__ddiv_rn(__dmul_rn(fy, y), zg), cy);
    const double un = __ddiv_rn(u, img_w), vn = __ddiv_rn(v, img_h);
    const bool ok = (un >= 0.0) & (un < 1.0) & (vn >= 0.0) & (vn < 1.0) & (z > 0.0);
    if (uvn) uvn[2 * i] = un, uvn[2 * i + 1] = vn;
    if (pm) pm[2 * i] = ok ? __double2int_rn(__dmul_rn(un, 1000.0)) : 0, pm[2 * i + 1] = ok ? __double2int_rn(__dmul_rn(vn, 1000.0)) : 0;
    if (valid) valid[i] = ok ? 1 : 0;
  }
}

// rigid_body_segmentation's O(T N^2) loop (single_object_movement_engine_coord.py:49-77): for every track pair (i<j)
// sum over t of |d_t - d_{t-1}| where that change exceeds `smoothing`, d_t = Euclidean distance at frame t (float64,
// accumulated in frame order like the reference).  One thread per pair; the tracks (T*N*12 B) stay in L2.
__global__ void __launch_bounds__(256) k_track_cumulative_loss(const float* __restrict__ tracks, int T, int N, double smoothing,
                                                               double* __restrict__ out) {
  const int64_t pidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t npairs = (int64_t)N * (N - 1) / 2;
  if (pidx >= npairs) return;
  // condensed index (scipy squareform order) -> (i, j), i < j
  int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)pidx)) * 0.5);
  while ((int64_t)i * (2 * N - i - 1) / 2 > pidx) --i;
  while ((int64_t)(i + 1) * (2 * N - i - 2) / 2 <= pidx) ++i;
  const int j = (int)(pidx - (int64_t)i * (2 * N - i - 1) / 2) + i + 1;
  double acc = 0.0, prev = 0.0;
  for (int t = 0; t < T; ++t) {
    const float* a = tracks + 3 * ((int64_t)t * N + i);
    const float* b = tracks + 3 * ((int64_t)t * N + j);
    const double dx = (double)a[0] - (double)b[0], dy = (double)a[1] - (double)b[1], dz = (double)a[2] - (double)b[2];
    const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    if (t > 0) {
      const double ch = fabs(d - prev);
      if (ch > smoothing) acc = __dadd_rn(acc, ch);
    }
    prev = d;
  }
  out[pidx] = acc;
}

__global__ void k_track_pair_displacement(const double* __restrict__ world, const double* __restrict__ w2c, int T, int N,
                                          const int32_t* __restrict__ triples, int64_t M, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int f1 = triples[3 * i], f2 = triples[3 * i + 1], pt = triples[3 * i + 2];
  const double* a = world + 3 * ((int64_t)f1 * N + pt);
  const double* b = world + 3 * ((int64_t)f2 * N + pt);
  const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  const double* w1 = w2c + (int64_t)f1 * 16;
  double m1[16], m2[16], c1[16], c2[16];
  for (int k = 0; k < 16; ++k) m1[k] = w1[k], m2[k] = w2c[(int64_t)f2 * 16 + k];
  inverse4(m1, c1);
  inverse4(m2, c2);
  const double ex = c2[3] - c1[3], ey = c2[7] - c1[7], ez = c2[11] - c1[11];
  out[5 * i + 0] = sqrt(__fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx))));
  // w2c[f1] @ [d, 0] is a (4,4)@(4,) product: same pairwise order as the single-column case
  for (int r = 0; r < 3; ++r)
    out[5 * i + 1 + r] = dot4_single(w1[4 * r], dx, w1[4 * r + 1], dy, w1[4 * r + 2], dz, w1[4 * r + 3], 0.0);
  out[5 * i + 4] = sqrt(__fma_rn(ez, ez, __fma_rn(ey, ey, __dmul_rn(ex, ex))));
}

// All i<j pairs of each selected track's visible frames (single_object_movement_engine_coord.py:484-498): the reference
// builds the pair list in a Python comprehension and calls np.linalg.norm(points2 - points1, axis=1), which sums the three
// squares as (dx*dx + dy*dy) + dz*dz with individually rounded products (NumPy's add.reduce over the short axis; no FMA).
// grid.y = track, grid.x covers that track's pairs; k -> (i, j) in lexicographic order.
__global__ void __launch_bounds__(256) k_track_all_pair_distances(const double* __restrict__ world, int N,
                                                                  const int32_t* __restrict__ track_ids,
                                                                  const int32_t* __restrict__ vis_frames,
                                                                  const int64_t* __restrict__ frame_offsets,
                                                                  const int64_t* __restrict__ pair_offsets,
                                                                  double* __restrict__ dist) {
  const int t = blockIdx.y;
  const int64_t fo = frame_offsets[t];
  const int64_t V = frame_offsets[t + 1] - fo;
  const int64_t P = V * (V - 1) / 2;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  int64_t i = (int64_t)((2.0 * V - 1.0 - sqrt((2.0 * V - 1.0) * (2.0 * V - 1.0) - 8.0 * (double)k)) * 0.5);
  if (i < 0) i = 0;
  if (i > V - 2) i = V - 2;
  while (i * (2 * V - i - 1) / 2 > k) --i;
  while ((i + 1) * (2 * V - i - 2) / 2 <= k) ++i;
  const int64_t j = k - i * (2 * V - i - 1) / 2 + i + 1;
  const int pt = track_ids[t];
  const double* a = world + 3 * ((int64_t)vis_frames[fo + i] * N + pt);
  const double* b = world + 3 * ((int64_t)vis_frames[fo + j] * N + pt);
  const double dx = __dsub_rn(b[0], a[0]), dy = __dsub_rn(b[1], a[1]), dz = __dsub_rn(b[2], a[2]);
  dist[pair_offsets[t] + k] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// ================================================================================================
// dot annotation: filled circles (OpenCV's midpoint walk) + glyph masks, one CTA per image's run of marks
// ================================================================================================
static_assert(sizeof(msp_mark) == 32, "msp_mark layout");

__global__ void __launch_bounds__(256) k_draw_marks(uint8_t* __restrict__ images, int B, int H, int W,
                                                    const msp_mark* __restrict__ marks, int64_t M,
                                                    const uint32_t* __restrict__ glyph_rows,
                                                    const int32_t* __restrict__ glyph_meta, int G) {
  const int64_t head = blockIdx.x;
  const int img = marks[head].image;
  if (head > 0 && marks[head - 1].image == img) return;      // only the first mark of a run paints the run, in order
  if (img < 0 || img >= B) return;
  __shared__ int s_hw[MSP_MAX_MARK_RADIUS + 1];               // half-width of the span on rows y +- k
  uint8_t* base = images + (int64_t)img * H * W * 3;
  for (int64_t j = head; j < M && marks[j].image == img; ++j) {
    const msp_mark m = marks[j];
    const int r = m.radius;
    if (r >= 0 && r <= MSP_MAX_MARK_RADIUS) {
      if (threadIdx.x == 0) {
        // cv::Circle (fill): (dx, dy) walks from (r, 0) while dx >= dy; rows +-dy get +-dx, rows +-dx get +-dy
        int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
        while (dx >= dy) {
          s_hw[dy] = dx;                                      // dx only shrinks: the first visit of row dy is the widest
          if (dx != dy) s_hw[dx] = dy;                        // dy only grows: the last visit of row dx wins
          ++dy;
          err += plus;
          plus += 2;
          if (err > 0) { err -= minus; --dx; minus -= 2; }
        }
      }
      __syncthreads();
      const int side = 2 * r + 1;
      for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
        const int ky = t / side - r, kx = t % side - r;
        if (abs(kx) > s_hw[abs(ky)]) continue;
        const int px = m.x + kx, py = m.y + ky;
        if (px < 0 || px >= W || py < 0 || py >= H) continue;
        uint8_t* q = base + ((int64_t)py * W + px) * 3;
        q[0] = m.bgr[0], q[1] = m.bgr[1], q[2] = m.bgr[2];
      }
      __syncthreads();
    }
    if (m.glyph >= 0 && m.glyph < G) {
      const int ox = glyph_meta[4 * m.glyph], oy = glyph_meta[4 * m.glyph + 1];
      const int gw = glyph_meta[4 * m.glyph + 2], gh = glyph_meta[4 * m.glyph + 3];
      for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
        const int row = t >> 5, col = t & 31;
        if (row >= gh || col >= gw || !((glyph_rows[32 * m.glyph + row] >> col) & 1u)) continue;
        const int px = m.org_x + ox + col, py = m.org_y + oy + row;
        if (px < 0 || px >= W || py < 0 || py >= H) continue;
        uint8_t* q = base + ((int64_t)py * W + px) * 3;
        q[0] = m.bgr[0], q[1] = m.bgr[1], q[2] = m.bgr[2];
      }
      __syncthreads();
    }
  }
}

// ================================================================================================
// C ABI
// ================================================================================================
static bool k_is_std(const double* K) {
  return K[1] == 0.0 && K[3] == 0.0 && K[4] == 0.0 && K[7] == 0.0 && K[8] == 0.0 && K[9] == 0.0 && K[10] == 1.0 &&
         K[11] == 0.0;
}
static Intrinsics make_intr(const double* K16) {
  Intrinsics k;
  memcpy(k.k, K16, sizeof(double) * 12);
  return k;
}
static int make_geom(const msp_camera* c, ImageGeom* g) {
  REQUIRE(c != nullptr, "host_cam is NULL");
  REQUIRE(c->image_h > 0 && c->image_w > 0 && c->depth_h > 0 && c->depth_w > 0, "non-positive image size");
  g->img_w = (double)c->image_w;
  g->img_h = (double)c->image_h;
  g->sx = (double)c->depth_w / (double)c->image_w;  // info_handler.py:356
  g->sy = (double)c->depth_h / (double)c->image_h;  // info_handler.py:357
  g->depth_scale = c->depth_scale;
  g->depth_w = c->depth_w;
  g->depth_h = c->depth_h;
  g->image_w = c->image_w;
  g->image_h = c->image_h;
  return MSP_OK;
}

extern "C" int msp_version(void) { return MSP_VERSION; }
extern "C" int msp_debug_set_cull_stats(unsigned long long* dev_counters2) {
  g_cull_stats = dev_counters2;
  return MSP_OK;
}
extern "C" const char* msp_last_error(void) { return g_err; }

extern "C" int msp_pose_prepare(const double* c2w, const double* axis_align, int32_t F, double* E_out, double* w2c_out,
                                double* yawpitch_out, int32_t* valid_out, void* stream) {
  REQUIRE(F >= 0, "F < 0");
  if (F == 0) return MSP_OK;
  REQUIRE(c2w && E_out && w2c_out && yawpitch_out && valid_out, "msp_pose_prepare: NULL pointer");
  k_pose_prepare<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c2w, axis_align, F, E_out, w2c_out, yawpitch_out,
                                                                     valid_out);
  return check_launch("msp_pose_prepare");
}

// MSP_DISABLE_CULL=1 forces the reference-order kernel for every point (A/B switch used by tests and profiling).
static bool cull_enabled() {
  const char* e = getenv("MSP_DISABLE_CULL");
  return !(e != nullptr && e[0] == '1');
}

// MSP_CULL_V9=0 selects the round-1 kernel (msp_cull.cuh) for the visibility fast path (A/B switch for tests and profiling).
static bool cull9_enabled() {
  const char* e = getenv("MSP_CULL_V9");
  return !(e != nullptr && e[0] == '0');
}

// Keep the stream-ordered workspace of K1 / K2 in the device's default memory pool between calls (the default release threshold
// of 0 hands it back to the driver at every synchronisation, and the next call pays a real allocation).
static void keep_workspace_pool() {
  static thread_local int pool_ready_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (pool_ready_dev != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_ready_dev = dev;
  }
}

// K1 second generation (msp_cull9.cuh): prepare pass into a stream-ordered workspace, then the main kernel.
static int launch_cull9(const FrameLoopParams& p, int dt, cudaStream_t s) {
  const unsigned tiles = (unsigned)((p.n + kC9TilePoints - 1) / kC9TilePoints);
  const dim3 grid(tiles, (unsigned)((p.F + kC9Frames - 1) / kC9Frames));
  REQUIRE(grid.y <= 65535u, "too many frames in one call");
  CullTilePrep* prep = nullptr;
  keep_workspace_pool();
  if (cudaMallocAsync(reinterpret_cast<void**>(&prep), sizeof(CullTilePrep) * (size_t)tiles, s) != cudaSuccess)
    return check_launch("msp_project_visibility(workspace)");
  const bool count = p.counts != nullptr;
  Cull9Params q{};
  CullParams& c = q.p;
  c.pts = p.pts, c.n = p.n, c.w2c = p.w2c, c.F = p.F, c.depth = p.depth, c.zbuf = nullptr;
  c.depth_stride = p.depth_stride, c.mask = p.mask, c.mask_stride = p.mask_stride, c.K = p.K, c.g = p.g;
  c.stats = g_cull_stats;
  c.inst = p.inst, c.counts = p.counts, c.n_objects = p.n_objects;
  q.prep = prep;
  const float fdw = (float)p.g.depth_w, fdh = (float)p.g.depth_h;
  q.nfdw = -fdw, q.nfdh = -fdh, q.fdw = fdw, q.fscale = (float)p.g.depth_scale;
  q.xclip = (fdw - 1.0f) + 8388608.0f, q.yclip = fdh - 1.0f;
  int rc = MSP_OK;
#define MSP_C9_LAUNCH(DT, CNT)                                                                                      \
  {                                                                                                                 \
    constexpr int smem = CNT ? (int)sizeof(CullTilePrep) : (int)offsetof(CullTilePrep, pt_slot);                    \
    k_cull_prepare<DT><<<tiles, kC9Threads, 0, s>>>(p.pts, p.n, count ? p.inst : nullptr, p.n_objects, prep);       \
    k_frame_cull9<DT, CNT><<<grid, kC9Threads, smem, s>>>(q);                                                       \
  }
#define MSP_C9(DT)                        \
  if (count) MSP_C9_LAUNCH(DT, true)      \
  else MSP_C9_LAUNCH(DT, false)
  switch (dt) {
    case 0: MSP_C9(0) break;
    case 1: MSP_C9(1) break;
    default: MSP_C9(2) break;
  }
#undef MSP_C9
#undef MSP_C9_LAUNCH
  rc = check_launch("k_frame_cull9");
  if (cudaFreeAsync(prep, s) != cudaSuccess && rc == MSP_OK) rc = check_launch("msp_project_visibility(workspace free)");
  return rc;
}

// K2 second generation (k_zbuf_cull9): prepare pass + packed projection + batched float64 splat.
static int launch_zbuf9(const FrameLoopParams& p, int dt, cudaStream_t s) {
  const unsigned tiles = (unsigned)((p.n + kC9TilePoints - 1) / kC9TilePoints);
  const dim3 grid(tiles, (unsigned)((p.F + kZ9Frames - 1) / kZ9Frames));
  REQUIRE(grid.y <= 65535u, "too many frames in one call");
  CullTilePrep* prep = nullptr;
  keep_workspace_pool();
  if (cudaMallocAsync(reinterpret_cast<void**>(&prep), sizeof(CullTilePrep) * (size_t)tiles, s) != cudaSuccess)
    return check_launch("msp_zbuffer_render(workspace)");
  Cull9Params q{};
  CullParams& c = q.p;
  c.pts = p.pts, c.n = p.n, c.w2c = p.w2c, c.F = p.F, c.depth = nullptr, c.zbuf = p.zbuf;
  c.depth_stride = p.depth_stride, c.K = p.K, c.g = p.g;
  q.prep = prep;
  const float fdw = (float)p.g.depth_w, fdh = (float)p.g.depth_h;
  q.nfdw = -fdw, q.nfdh = -fdh, q.fdw = fdw, q.fscale = (float)p.g.depth_scale;
  q.xclip = (fdw - 1.0f) + 8388608.0f, q.yclip = fdh - 1.0f;
  constexpr int smem = (int)offsetof(CullTilePrep, pt_slot);
  switch (dt) {
    case 0:
      k_cull_prepare<0><<<tiles, kC9Threads, 0, s>>>(p.pts, p.n, nullptr, 0, prep);
      k_zbuf_cull9<0><<<grid, kC9Threads, smem, s>>>(q);
      break;
    case 1:
      k_cull_prepare<1><<<tiles, kC9Threads, 0, s>>>(p.pts, p.n, nullptr, 0, prep);
      k_zbuf_cull9<1><<<grid, kC9Threads, smem, s>>>(q);
      break;
    default:
      k_cull_prepare<2><<<tiles, kC9Threads, 0, s>>>(p.pts, p.n, nullptr, 0, prep);
      k_zbuf_cull9<2><<<grid, kC9Threads, smem, s>>>(q);
      break;
  }
  int rc = check_launch("k_zbuf_cull9");
  if (cudaFreeAsync(prep, s) != cudaSuccess && rc == MSP_OK) rc = check_launch("msp_zbuffer_render(workspace free)");
  return rc;
}

template <int MODE>
static int launch_frame_loop(const FrameLoopParams& p, int dt, bool kstd, bool mat, cudaStream_t s) {
  if (MODE == MODE_ZBUF && kstd && cull_enabled() && cull9_enabled() && p.depth_stride < (int64_t)kZ9NoPixel) return launch_zbuf9(p, dt, s);
  // the second-generation kernel gathers depth pixels as aligned 32-bit words (cp.async): every map must start on a 4-byte boundary
  const bool depth_words_ok = (reinterpret_cast<uintptr_t>(p.depth) & 3u) == 0 && (p.depth_stride & 1) == 0;
  if (MODE == MODE_VIS && kstd && !mat && cull_enabled() && cull9_enabled() && depth_words_ok) return launch_cull9(p, dt, s);
  if (kstd && !mat && cull_enabled()) {
    // fast path: float32 conservative cull + exact float64 on compacted survivors (msp_cull.cuh)
    const dim3 grid((unsigned)((p.n + kCullTilePoints - 1) / kCullTilePoints),
                    (unsigned)((p.F + kCullFramesPerCta - 1) / kCullFramesPerCta));
    REQUIRE(grid.y <= 65535u, "too many frames in one call");
    CullParams c{};
    c.pts = p.pts, c.n = p.n, c.w2c = p.w2c, c.F = p.F, c.depth = p.depth, c.zbuf = p.zbuf;
    c.depth_stride = p.depth_stride, c.mask = p.mask, c.mask_stride = p.mask_stride, c.K = p.K, c.g = p.g;
    c.stats = g_cull_stats;
    c.fdw = (float)p.g.depth_w, c.fdh = (float)p.g.depth_h, c.fdw1 = c.fdw - 1.0f, c.fdh1 = c.fdh - 1.0f;
    c.fscale = (float)p.g.depth_scale;
    c.inst = p.inst, c.counts = p.counts, c.n_objects = p.n_objects;
    const bool count = (MODE == MODE_VIS) && p.counts != nullptr;
#define MSP_CULL_LAUNCH(DT, CNT)                                                                             \
  {                                                                                                          \
    constexpr int smem = kCullTilePoints * PtTraits<DT>::kBytes;                                             \
    /* the attribute is per device and this library may drive several: set it on every call (cheap, thread-safe) */ \
    if (cudaFuncSetAttribute(k_frame_cull<DT, MODE, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) \
      return check_launch("k_frame_cull(cudaFuncSetAttribute)");                                            \
    k_frame_cull<DT, MODE, CNT><<<grid, kCullThreads, smem, s>>>(c);                                         \
  }
#define MSP_CULL(DT)                                      \
  if (MODE == MODE_VIS && count) MSP_CULL_LAUNCH(DT, true) \
  else MSP_CULL_LAUNCH(DT, false)
    switch (dt) {
      case 0: MSP_CULL(0) break;
      case 1: MSP_CULL(1) break;
      default: MSP_CULL(2) break;
    }
#undef MSP_CULL_LAUNCH
#undef MSP_CULL
    return MSP_OK;
  }
  REQUIRE(p.counts == nullptr, "internal: fused counting needs the fast path");
  const dim3 grid((unsigned)((p.n + kTilePoints - 1) / kTilePoints), (unsigned)((p.F + kFramesPerCta - 1) / kFramesPerCta));
  REQUIRE(grid.y <= 65535u, "too many frames in one call (max 65535*32)");
#define MSP_LAUNCH(DT, KS, MT) k_frame_loop<DT, KS, MODE, MT><<<grid, kThreads, 0, s>>>(p)
#define MSP_DISPATCH_DT(KS, MT)       \
  switch (dt) {                       \
    case 0: MSP_LAUNCH(0, KS, MT); break; \
    case 1: MSP_LAUNCH(1, KS, MT); break; \
    default: MSP_LAUNCH(2, KS, MT); break; \
  }
  if (MODE == MODE_VIS && mat) {
    if (kstd) { MSP_DISPATCH_DT(true, true) } else { MSP_DISPATCH_DT(false, true) }
  } else {
    if (kstd) { MSP_DISPATCH_DT(true, false) } else { MSP_DISPATCH_DT(false, false) }
  }
#undef MSP_DISPATCH_DT
#undef MSP_LAUNCH
  return MSP_OK;
}

extern "C" int msp_project_visibility(const void* pts, int64_t n, int32_t pts_dtype, const double* K16,
                                      const double* w2c, int32_t F, const uint16_t* depth, const msp_camera* host_cam,
                                      uint32_t* mask_bits, int64_t mask_stride_words, double* xy_out, double* z_out,
                                      void* stream) {
  REQUIRE(n >= 0 && F >= 0, "negative size");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  REQUIRE((xy_out == nullptr) == (z_out == nullptr), "xy_out and z_out must both be given or both be NULL");
  if (n == 0 || F == 0) return MSP_OK;
  REQUIRE(pts && K16 && w2c && depth && mask_bits, "msp_project_visibility: NULL pointer");
  REQUIRE(aligned16(pts) && aligned16(mask_bits), "pts and mask_bits must be 16-byte aligned");
  REQUIRE(mask_stride_words % 4 == 0 && mask_stride_words >= (n + 31) / 32, "mask_stride_words must be a multiple of 4 and >= ceil(n/32)");
  REQUIRE(xy_out == nullptr || aligned16(xy_out), "xy_out must be 16-byte aligned");
  FrameLoopParams p{};
  if (int rc = make_geom(host_cam, &p.g)) return rc;
  p.pts = pts, p.n = n, p.w2c = w2c, p.F = F, p.depth = depth, p.zbuf = nullptr;
  p.depth_stride = (int64_t)host_cam->depth_h * host_cam->depth_w;
  p.mask = mask_bits, p.mask_stride = mask_stride_words, p.xy = xy_out, p.z = z_out;
  p.K = make_intr(K16);
  if (int rc = launch_frame_loop<MODE_VIS>(p, pts_dtype, k_is_std(K16), xy_out != nullptr, (cudaStream_t)stream)) return rc;
  return check_launch("msp_project_visibility");
}

extern "C" int msp_project_visibility_objects(const void* pts, int64_t n, int32_t pts_dtype, const double* K16,
                                              const double* w2c, int32_t F, const uint16_t* depth,
                                              const msp_camera* host_cam, uint32_t* mask_bits, int64_t mask_stride_words,
                                              const uint16_t* instance_ids, int32_t n_objects, uint32_t* counts,
                                              void* stream) {
  REQUIRE(n >= 0 && F >= 0 && n_objects >= 0, "negative size");
  REQUIRE(n_objects <= 8191, "n_objects must be <= 8191");
  if (F == 0) return MSP_OK;
  REQUIRE(n_objects == 0 || (instance_ids && counts), "msp_project_visibility_objects: NULL instance_ids / counts");
  // the second-generation kernel counts any number of objects in the same pass (per-tile local id table); the round-1
  // kernel (MSP_CULL_V9=0) fuses up to kCullMaxObjects and otherwise needs the second launch below
  const bool depth_words_ok = host_cam != nullptr && (reinterpret_cast<uintptr_t>(depth) & 3u) == 0 &&
                              (((int64_t)host_cam->depth_h * host_cam->depth_w) & 1) == 0;
  const bool fused = n_objects > 0 && ((cull9_enabled() && depth_words_ok) || n_objects <= kCullMaxObjects) && n > 0 && K16 != nullptr &&
                     k_is_std(K16) && cull_enabled();
  if (!fused) {  // same result in two launches
    if (int rc = msp_project_visibility(pts, n, pts_dtype, K16, w2c, F, depth, host_cam, mask_bits, mask_stride_words,
                                        nullptr, nullptr, stream)) return rc;
    return msp_object_counts(mask_bits, n, mask_stride_words, F, instance_ids, n_objects, counts, stream);
  }
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  REQUIRE(pts && w2c && depth && mask_bits, "msp_project_visibility_objects: NULL pointer");
  REQUIRE(aligned16(pts) && aligned16(mask_bits), "pts and mask_bits must be 16-byte aligned");
  REQUIRE(mask_stride_words % 4 == 0 && mask_stride_words >= (n + 31) / 32, "mask_stride_words must be a multiple of 4 and >= ceil(n/32)");
  FrameLoopParams p{};
  if (int rc = make_geom(host_cam, &p.g)) return rc;
  p.pts = pts, p.n = n, p.w2c = w2c, p.F = F, p.depth = depth, p.zbuf = nullptr;
  p.depth_stride = (int64_t)host_cam->depth_h * host_cam->depth_w;
  p.mask = mask_bits, p.mask_stride = mask_stride_words;
  p.K = make_intr(K16);
  p.inst = instance_ids, p.counts = counts, p.n_objects = n_objects;
  if (cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)F * (size_t)n_objects, (cudaStream_t)stream) != cudaSuccess)
    return check_launch("msp_project_visibility_objects(memset)");
  if (int rc = launch_frame_loop<MODE_VIS>(p, pts_dtype, true, false, (cudaStream_t)stream)) return rc;
  return check_launch("msp_project_visibility_objects");
}

extern "C" int msp_zbuffer_render(const void* pts, int64_t n, int32_t pts_dtype, const double* K16, const double* w2c,
                                  int32_t F, const msp_camera* host_cam, float* zbuf, void* stream) {
  REQUIRE(n >= 0 && F >= 0, "negative size");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  if (F == 0) return MSP_OK;
  REQUIRE(K16 && w2c && zbuf && host_cam, "msp_zbuffer_render: NULL pointer");
  FrameLoopParams p{};
  if (int rc = make_geom(host_cam, &p.g)) return rc;
  p.depth_stride = (int64_t)host_cam->depth_h * host_cam->depth_w;
  const int64_t total = p.depth_stride * F;
  k_fill_u32<<<(unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<uint32_t*>(zbuf), total, 0x7f800000u);
  if (n > 0) {
    REQUIRE(pts != nullptr && aligned16(pts), "pts must be non-NULL and 16-byte aligned");
    p.pts = pts, p.n = n, p.w2c = w2c, p.F = F, p.zbuf = zbuf;
    p.K = make_intr(K16);
    if (int rc = launch_frame_loop<MODE_ZBUF>(p, pts_dtype, k_is_std(K16), false, (cudaStream_t)stream)) return rc;
  }
  return check_launch("msp_zbuffer_render");
}

extern "C" int msp_depth_to_u16(const float* zbuf, int64_t count, int32_t margin_mm, double depth_scale, uint16_t* out,
                                void* stream) {
  REQUIRE(count >= 0, "negative count");
  if (count == 0) return MSP_OK;
  REQUIRE(zbuf && out, "msp_depth_to_u16: NULL pointer");
  REQUIRE(depth_scale > 0.0, "depth_scale must be positive");
  REQUIRE(aligned16(zbuf) && aligned16(out), "zbuf and out must be 16-byte aligned");
  REQUIRE(margin_mm >= -1 && margin_mm <= 60000, "margin_mm out of range");
  const int64_t work = (count + 7) / 8;
  k_depth_to_u16<<<(unsigned)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      zbuf, count, margin_mm, depth_scale, out);
  return check_launch("msp_depth_to_u16");
}

extern "C" int msp_pair_overlap(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, const int32_t* pairs,
                                int64_t P, uint32_t* inter, uint32_t* uni, double* overlap_pct, void* stream) {
  REQUIRE(n >= 0 && P >= 0, "negative size");
  if (P == 0) return MSP_OK;
  REQUIRE(mask_bits && pairs && inter && uni && overlap_pct, "msp_pair_overlap: NULL pointer");
  REQUIRE(aligned16(mask_bits) && mask_stride_words % 4 == 0 && mask_stride_words >= (n + 31) / 32, "mask layout");
  PairParams p{};
  p.mask = mask_bits, p.n = n, p.mask_stride = mask_stride_words, p.pairs = pairs, p.P = P;
  p.inter_out = inter, p.uni_out = uni, p.overlap_out = overlap_pct;
  k_pair<0, true, false><<<(unsigned)P, kPairThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("msp_pair_overlap");
}

extern "C" int msp_pair_annotate(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, const int32_t* pairs,
                                 int64_t P, const uint32_t* sel, int32_t sel_mode, const void* pts, int32_t pts_dtype,
                                 const double* K16, const double* E, const double* w2c, const double* yawpitch,
                                 const uint16_t* depth, const msp_camera* host_cam, msp_pair_record* records,
                                 void* stream) {
  REQUIRE(n >= 0 && P >= 0, "negative size");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  REQUIRE(sel_mode == 0 || sel_mode == 1, "sel_mode must be 0 or 1");
  REQUIRE(P < (1ll << 31), "P too large for one launch");
  if (P == 0) return MSP_OK;
  REQUIRE(mask_bits && pairs && sel && pts && K16 && E && w2c && yawpitch && depth && records, "msp_pair_annotate: NULL pointer");
  REQUIRE(aligned16(mask_bits) && mask_stride_words % 4 == 0 && mask_stride_words >= (n + 31) / 32, "mask layout");
  REQUIRE(aligned16(pts) && aligned16(records), "pts and records must be 16-byte aligned");
  PairParams p{};
  if (int rc = make_geom(host_cam, &p.g)) return rc;
  p.mask = mask_bits, p.n = n, p.mask_stride = mask_stride_words, p.pairs = pairs, p.P = P, p.sel = sel;
  p.sel_mode = sel_mode, p.pts = pts, p.E = E, p.w2c = w2c, p.yawpitch = yawpitch, p.depth = depth;
  p.depth_stride = (int64_t)host_cam->depth_h * host_cam->depth_w;
  p.K = make_intr(K16);
  p.rec = records;
  const bool ks = k_is_std(K16);
  cudaStream_t s = (cudaStream_t)stream;
#define MSP_PAIR(DT)                                                            \
  if (ks) k_pair<DT, true, true><<<(unsigned)P, kPairThreads, 0, s>>>(p);      \
  else k_pair<DT, false, true><<<(unsigned)P, kPairThreads, 0, s>>>(p)
  switch (pts_dtype) {
    case 0: MSP_PAIR(0); break;
    case 1: MSP_PAIR(1); break;
    default: MSP_PAIR(2); break;
  }
#undef MSP_PAIR
  return check_launch("msp_pair_annotate");
}

extern "C" int msp_project_points(const void* pts, int64_t n, int32_t pts_dtype, const double* K16, const double* w2c_one,
                                  double* uvz, void* stream) {
  REQUIRE(n >= 0, "negative size");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  if (n == 0) return MSP_OK;
  REQUIRE(pts && K16 && w2c_one && uvz, "msp_project_points: NULL pointer");
  REQUIRE(aligned16(pts), "pts must be 16-byte aligned");
  AnnotParams p{};
  msp_camera cam{1, 1, 1, 1, 1.0};
  if (int rc = make_geom(&cam, &p.g)) return rc;
  p.pts = pts, p.n = n, p.point_ids = nullptr, p.frame_ids = nullptr, p.M = n, p.w2c = w2c_one, p.depth = nullptr;
  p.K = make_intr(K16);
  p.uvz = uvz;
  p.batch_order = (n >= 2) ? 1 : 0;   // NumPy multiplies a single column differently (see msp_device.cuh)
  const bool ks = k_is_std(K16);
  const unsigned grid = (unsigned)((n + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
#define MSP_PRJ(DT)                                        \
  if (ks) k_annotate<DT, true><<<grid, 128, 0, s>>>(p);   \
  else k_annotate<DT, false><<<grid, 128, 0, s>>>(p)
  switch (pts_dtype) {
    case 0: MSP_PRJ(0); break;
    case 1: MSP_PRJ(1); break;
    default: MSP_PRJ(2); break;
  }
#undef MSP_PRJ
  return check_launch("msp_project_points");
}

extern "C" int msp_annotate_points(const void* pts, int64_t n, int32_t pts_dtype, const int32_t* point_ids,
                                   const int32_t* frame_ids, int64_t M, const double* K16, const double* w2c,
                                   const uint16_t* depth, const msp_camera* host_cam, double* uvz, int32_t* permille,
                                   int32_t* depth_mm, int32_t* pix, uint8_t* visible, void* stream) {
  REQUIRE(M >= 0 && n >= 0, "negative size");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  if (M == 0) return MSP_OK;
  REQUIRE(pts && point_ids && frame_ids && K16 && w2c, "msp_annotate_points: NULL pointer");
  REQUIRE(aligned16(pts), "pts must be 16-byte aligned");
  AnnotParams p{};
  if (int rc = make_geom(host_cam, &p.g)) return rc;
  p.pts = pts, p.n = n, p.point_ids = point_ids, p.frame_ids = frame_ids, p.M = M, p.w2c = w2c, p.depth = depth;
  p.depth_stride = (int64_t)host_cam->depth_h * host_cam->depth_w;
  p.K = make_intr(K16);
  p.uvz = uvz, p.permille = permille, p.depth_mm = depth_mm, p.pix = pix, p.visible = visible;
  const bool ks = k_is_std(K16);
  const unsigned grid = (unsigned)((M + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
#define MSP_ANN(DT)                                        \
  if (ks) k_annotate<DT, true><<<grid, 128, 0, s>>>(p);   \
  else k_annotate<DT, false><<<grid, 128, 0, s>>>(p)
  switch (pts_dtype) {
    case 0: MSP_ANN(0); break;
    case 1: MSP_ANN(1); break;
    default: MSP_ANN(2); break;
  }
#undef MSP_ANN
  return check_launch("msp_annotate_points");
}

extern "C" int msp_check_visibility(const double* xy, const double* z, int64_t n, const uint16_t* depth,
                                    const msp_camera* host_cam, int32_t mode, uint8_t* out, void* stream) {
  REQUIRE(n >= 0, "negative size");
  REQUIRE(mode >= 1 && mode <= 3, "mode must be 1 (bounds), 2 (depth) or 3 (both)");
  if (n == 0) return MSP_OK;
  REQUIRE(xy && out, "msp_check_visibility: NULL pointer");
  REQUIRE(!(mode & 2) || (depth && z), "msp_check_visibility: depth test needs depth and z");
  ImageGeom g;
  if (int rc = make_geom(host_cam, &g)) return rc;
  k_check_visibility<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xy, z, n, depth, g, mode, out);
  return check_launch("msp_check_visibility");
}

extern "C" int msp_object_counts(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, int32_t F,
                                 const uint16_t* instance_ids, int32_t n_objects, uint32_t* counts, void* stream) {
  REQUIRE(n >= 0 && F >= 0 && n_objects >= 0, "negative size");
  REQUIRE(n_objects <= 8191, "n_objects must be <= 8191");
  if (F == 0 || n_objects == 0) return MSP_OK;
  REQUIRE(mask_bits && instance_ids && counts, "msp_object_counts: NULL pointer");
  REQUIRE(mask_stride_words >= (n + 31) / 32, "mask_stride_words too small");
  k_object_counts<<<(unsigned)F, 256, (size_t)(n_objects + 1) * sizeof(uint32_t), (cudaStream_t)stream>>>(
      mask_bits, n, mask_stride_words, instance_ids, n_objects, counts);
  return check_launch("msp_object_counts");
}

extern "C" int msp_object_extents(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, int32_t F,
                                  const void* pts, int32_t pts_dtype, const uint16_t* instance_ids, int32_t n_objects,
                                  double* extents, void* stream) {
  REQUIRE(n >= 0 && F >= 0 && n_objects >= 0, "negative size");
  REQUIRE(n_objects <= 768, "n_objects must be <= 768");
  REQUIRE(pts_dtype >= 0 && pts_dtype <= 2, "pts_dtype must be 0, 1 or 2");
  if (F == 0 || n_objects == 0) return MSP_OK;
  REQUIRE(mask_bits && pts && instance_ids && extents, "msp_object_extents: NULL pointer");
  REQUIRE(mask_stride_words >= (n + 31) / 32, "mask_stride_words too small");
  const size_t smem = (size_t)n_objects * 6 * sizeof(long long);
  cudaStream_t s = (cudaStream_t)stream;
#define MSP_EXT(DT) k_object_extents<DT><<<(unsigned)F, 256, smem, s>>>(mask_bits, n, mask_stride_words, pts, instance_ids, n_objects, extents)
  switch (pts_dtype) {
    case 0: MSP_EXT(0); break;
    case 1: MSP_EXT(1); break;
    default: MSP_EXT(2); break;
  }
#undef MSP_EXT
  return check_launch("msp_object_extents");
}

extern "C" int msp_track_project(const float* tracks, const double* w2c, const double* fxfycxcy, int32_t T, int32_t N,
                                 int32_t image_h, int32_t image_w, double* world_out, double* uvn_out,
                                 int32_t* permille_out, uint8_t* valid_out, void* stream) {
  REQUIRE(T >= 0 && N >= 0, "negative size");
  REQUIRE(image_h > 0 && image_w > 0, "non-positive image size");
  if (T == 0 || N == 0) return MSP_OK;
  REQUIRE(tracks && w2c && fxfycxcy, "msp_track_project: NULL pointer");
  k_track_project<<<(unsigned)T, 128, 0, (cudaStream_t)stream>>>(tracks, w2c, fxfycxcy, T, N, (double)image_h,
                                                                 (double)image_w, world_out, uvn_out, permille_out,
                                                                 valid_out);
  return check_launch("msp_track_project");
}

extern "C" int msp_track_cumulative_loss(const float* tracks, int32_t T, int32_t N, double smoothing, double* condensed_out,
                                         void* stream) {
  REQUIRE(T >= 0 && N >= 0, "negative size");
  if (N < 2) return MSP_OK;
  REQUIRE(tracks && condensed_out, "msp_track_cumulative_loss: NULL pointer");
  const int64_t npairs = (int64_t)N * (N - 1) / 2;
  REQUIRE(npairs < (1ll << 31) * 256, "too many track pairs");
  k_track_cumulative_loss<<<(unsigned)((npairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tracks, T, N, smoothing,
                                                                                               condensed_out);
  return check_launch("msp_track_cumulative_loss");
}

extern "C" int msp_track_pair_displacement(const double* world, const double* w2c, int32_t T, int32_t N,
                                           const int32_t* triples, int64_t M, double* out, void* stream) {
  REQUIRE(T >= 0 && N >= 0 && M >= 0, "negative size");
  if (M == 0) return MSP_OK;
  REQUIRE(world && w2c && triples && out, "msp_track_pair_displacement: NULL pointer");
  k_track_pair_displacement<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(world, w2c, T, N, triples,
                                                                                          M, out);
  return check_launch("msp_track_pair_displacement");
}

extern "C" int msp_track_all_pair_distances(const double* world, int32_t T, int32_t N, const int32_t* track_ids,
                                            int32_t n_tracks, const int32_t* vis_frames, const int64_t* frame_offsets,
                                            const int64_t* pair_offsets, int64_t max_pairs_per_track, double* dist,
                                            void* stream) {
  REQUIRE(T >= 0 && N >= 0 && n_tracks >= 0 && max_pairs_per_track >= 0, "negative size");
  if (n_tracks == 0 || max_pairs_per_track == 0) return MSP_OK;
  REQUIRE(world && track_ids && vis_frames && frame_offsets && pair_offsets && dist, "msp_track_all_pair_distances: NULL pointer");
  REQUIRE(n_tracks <= 65535, "at most 65535 tracks per call");
  REQUIRE(max_pairs_per_track < (1ll << 31) * 256, "too many pairs per track");
  const dim3 grid((unsigned)((max_pairs_per_track + 255) / 256), (unsigned)n_tracks);
  k_track_all_pair_distances<<<grid, 256, 0, (cudaStream_t)stream>>>(world, N, track_ids, vis_frames, frame_offsets,
                                                                     pair_offsets, dist);
  return check_launch("msp_track_all_pair_distances");
}

extern "C" int msp_draw_marks(uint8_t* images, int32_t B, int32_t H, int32_t W, const msp_mark* marks, int64_t M,
                              const uint32_t* glyph_rows, const int32_t* glyph_meta, int32_t G, void* stream) {
  REQUIRE(B >= 0 && H >= 0 && W >= 0 && M >= 0 && G >= 0, "negative size");
  if (M == 0 || B == 0 || H == 0 || W == 0) return MSP_OK;
  REQUIRE(images && marks, "msp_draw_marks: NULL pointer");
  REQUIRE(G == 0 || (glyph_rows && glyph_meta), "msp_draw_marks: glyph atlas missing");
  REQUIRE(M < (1ll << 31), "too many marks");
  k_draw_marks<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(images, B, H, W, marks, M, glyph_rows, glyph_meta, G);
  return check_launch("msp_draw_marks");
}

// ---- visibility table text (msp_tables.cuh)
static int json_common_checks(const uint32_t* mask_bits, int64_t n, int64_t stride, int32_t R) {
  REQUIRE(n >= 0 && R >= 0, "negative size");
  REQUIRE(n < (1ll << 31), "point ids must fit int32");
  REQUIRE(mask_bits != nullptr || n == 0 || R == 0, "mask_bits is NULL");
  REQUIRE(stride >= (n + 31) / 32, "mask_stride_words too small");
  return MSP_OK;
}

extern "C" int msp_visibility_json_points_size(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                               const int32_t* frame_rows, int32_t R, int64_t* row_bytes, void* stream) {
  if (int rc = json_common_checks(mask_bits, n, mask_stride_words, R)) return rc;
  if (R == 0) return MSP_OK;
  REQUIRE(row_bytes != nullptr, "msp_visibility_json_points_size: NULL pointer");
  k_json_point_lists<false><<<(unsigned)R, kJsonThreads, 0, (cudaStream_t)stream>>>(mask_bits, n, mask_stride_words, frame_rows,
                                                                                    nullptr, row_bytes, nullptr);
  return check_launch("msp_visibility_json_points_size");
}

extern "C" int msp_visibility_json_points_write(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                                const int32_t* frame_rows, int32_t R, const int64_t* offsets, uint8_t* text,
                                                void* stream) {
  if (int rc = json_common_checks(mask_bits, n, mask_stride_words, R)) return rc;
  if (R == 0) return MSP_OK;
  REQUIRE(offsets && text, "msp_visibility_json_points_write: NULL pointer");
  k_json_point_lists<true><<<(unsigned)R, kJsonThreads, 0, (cudaStream_t)stream>>>(mask_bits, n, mask_stride_words, frame_rows,
                                                                                   offsets, nullptr, text);
  return check_launch("msp_visibility_json_points_write");
}

extern "C" int msp_visibility_json_images_size(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                               const int32_t* frame_rows, int32_t R, const uint8_t* id_len,
                                               int64_t* point_bytes, void* stream) {
  if (int rc = json_common_checks(mask_bits, n, mask_stride_words, R)) return rc;
  if (n == 0) return MSP_OK;
  REQUIRE(point_bytes && (id_len || R == 0), "msp_visibility_json_images_size: NULL pointer");
  k_json_image_lists<false><<<(unsigned)((n + kJsonThreads - 1) / kJsonThreads), kJsonThreads, 0, (cudaStream_t)stream>>>(
      mask_bits, n, mask_stride_words, frame_rows, R, nullptr, 0, id_len, nullptr, point_bytes, nullptr);
  return check_launch("msp_visibility_json_images_size");
}

extern "C" int msp_visibility_json_images_write(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                                const int32_t* frame_rows, int32_t R, const uint8_t* ids, int32_t id_pitch,
                                                const uint8_t* id_len, const int64_t* offsets, uint8_t* text, void* stream) {
  if (int rc = json_common_checks(mask_bits, n, mask_stride_words, R)) return rc;
  if (n == 0) return MSP_OK;
  REQUIRE(offsets && text && ((ids && id_len && id_pitch > 0) || R == 0), "msp_visibility_json_images_write: NULL pointer");
  k_json_image_lists<true><<<(unsigned)((n + kJsonThreads - 1) / kJsonThreads), kJsonThreads, 0, (cudaStream_t)stream>>>(
      mask_bits, n, mask_stride_words, frame_rows, R, ids, id_pitch, id_len, offsets, nullptr, text);
  return check_launch("msp_visibility_json_images_write");
}

#include "msp_jpeg.inl"
