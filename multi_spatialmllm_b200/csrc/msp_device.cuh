// Device-side building blocks of the MultiSPA geometry hot path (sm_100a).
//
// Arithmetic contract.  Everything that decides an integer (mask bit, pixel index, per-mille value) is
// evaluated in float64 with explicitly rounded operations in the order the reference's NumPy/BLAS code
// uses: np.linalg.inv(E) once per frame, cam = Einv @ P^T as a k-sequential FMA chain (what OpenBLAS
// dgemm does for the 4-term dot products; verified bit-for-bit against numpy in this repo's tests),
// img = K @ cam likewise, then a true IEEE division (info_handler.py:57-69).  Explicit __fma_rn/__dmul_rn
// keep nvcc from re-associating or contracting differently.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace msp {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  // make the init visible to the async proxy before a bulk copy signals the barrier
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#ifndef MSP_GATHER_MODE
#define MSP_GATHER_MODE 0
#endif
// depth-map gather (read-only path).  MSP_GATHER_MODE selects the cache hints (tuning, profiles/tuning_r1.md)
__device__ __forceinline__ uint16_t ldg_u16_nc(const uint16_t* p) {
#if MSP_GATHER_MODE == 0
  return __ldg(p);
#else
  uint16_t v;
#if MSP_GATHER_MODE == 1
  asm volatile("ld.global.nc.L2::256B.u16 %0, [%1];" : "=h"(v) : "l"(p));
#elif MSP_GATHER_MODE == 2
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u16 %0, [%1];" : "=h"(v) : "l"(p));
#elif MSP_GATHER_MODE == 3
  asm volatile("ld.global.nc.L1::evict_last.L2::128B.u16 %0, [%1];" : "=h"(v) : "l"(p));
#else
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
#endif
  return v;
#endif
}

// ------------------------------------------------------------------------------------------------
// camera model
// ------------------------------------------------------------------------------------------------
struct Intrinsics {  // rows 0..2 of the reference's 4x4 K (info_handler.py:11, update_info_file_with_images.py:55-69)
  double k[12];
};
struct ImageGeom {
  double img_w, img_h;  // float64 copies of the integer sizes (compared against u, v: IH:342-343)
  double sx, sy;        // depth_w / image_w, depth_h / image_h (IH:356-357)
  double depth_scale;   // IH:368
  int32_t depth_w, depth_h;
  int32_t image_w, image_h;
};

// cam = (w2c rows 0..2) . [x y z 1], k-sequential FMA chain (IH:60)
__device__ __forceinline__ double row_dot_h(const double* __restrict__ r, double x, double y, double z) {
  return __dadd_rn(__fma_rn(r[2], z, __fma_rn(r[1], y, __dmul_rn(r[0], x))), r[3]);
}
__device__ __forceinline__ double row_dot4(const double* __restrict__ r, double a, double b, double c, double d) {
  return __fma_rn(r[3], d, __fma_rn(r[2], c, __fma_rn(r[1], b, __dmul_rn(r[0], a))));
}

// The same products when NumPy multiplies a 4x4 by a SINGLE column ((4,4)@(4,1), as every per-point call of the
// reference does: get_point_2d_coordinates_in_image, info_handler.py:291-305 -> IH:331,60,66).  That shape takes a
// different BLAS path whose 4-term dot product is (a0 b0 + a2 b2) + (a1 b1 + a3 b3) with individually rounded products
// and no FMA (verified bit-for-bit, 0 mismatches in 1 200 random dot products; tests/test_oracle_golden.py).
__device__ __forceinline__ double dot4_single(double a0, double b0, double a1, double b1, double a2, double b2, double a3,
                                              double b3) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a2, b2)), __dadd_rn(__dmul_rn(a1, b1), __dmul_rn(a3, b3)));
}
__device__ __forceinline__ double row_dot_h_single(const double* __restrict__ r, double x, double y, double z) {
  return dot4_single(r[0], x, r[1], y, r[2], z, r[3], 1.0);
}

struct Proj {
  double u, v, z;
};

// Exact projection of one point.  STD = the w2c bottom row is [0,0,0,1] and K is [fx 0 cx 0; 0 fy cy 0; 0 0 1 0],
// for which the zero terms of the BLAS dot products vanish exactly and cam.w == 1.
// SINGLE selects the one-column order above (per-point calls) instead of the N >= 2 gemm order (batch calls).
template <bool STD, bool SINGLE = false>
__device__ __forceinline__ Proj project_exact(const double* __restrict__ w /*16 or 12*/, const Intrinsics& K, double x,
                                              double y, double z) {
  Proj p;
  if (SINGLE) {
    const double X = row_dot_h_single(w + 0, x, y, z);
    const double Y = row_dot_h_single(w + 4, x, y, z);
    const double Z = row_dot_h_single(w + 8, x, y, z);
    const double W = row_dot_h_single(w + 12, x, y, z);
    p.z = Z;
    const double i0 = dot4_single(K.k[0], X, K.k[1], Y, K.k[2], Z, K.k[3], W);
    const double i1 = dot4_single(K.k[4], X, K.k[5], Y, K.k[6], Z, K.k[7], W);
    const double i2 = dot4_single(K.k[8], X, K.k[9], Y, K.k[10], Z, K.k[11], W);
    p.u = __ddiv_rn(i0, i2);
    p.v = __ddiv_rn(i1, i2);
    return p;
  }
  const double X = row_dot_h(w + 0, x, y, z);
  const double Y = row_dot_h(w + 4, x, y, z);
  const double Z = row_dot_h(w + 8, x, y, z);
  p.z = Z;  // points_depth = camera_coords[2] (IH:63)
  if (STD) {
    const double i0 = __fma_rn(K.k[2], Z, __dmul_rn(K.k[0], X));
    const double i1 = __fma_rn(K.k[6], Z, __dmul_rn(K.k[5], Y));
    p.u = __ddiv_rn(i0, Z);
    p.v = __ddiv_rn(i1, Z);
  } else {
    const double W = row_dot_h(w + 12, x, y, z);
    const double i0 = row_dot4(K.k + 0, X, Y, Z, W);
    const double i1 = row_dot4(K.k + 4, X, Y, Z, W);
    const double i2 = row_dot4(K.k + 8, X, Y, Z, W);
    p.u = __ddiv_rn(i0, i2);
    p.v = __ddiv_rn(i1, i2);
  }
  return p;
}

__device__ __forceinline__ bool in_image(const Proj& p, const ImageGeom& g) {  // IH:342-343
  return (p.u >= 0.0) & (p.u < g.img_w) & (p.v >= 0.0) & (p.v < g.img_h);
}

// IH:359-366 for a point already known to be inside the colour image (so the products are in [0, depth_w]).
__device__ __forceinline__ int depth_index(const Proj& p, const ImageGeom& g) {
  int dx = __double2int_rn(__dmul_rn(p.u, g.sx));  // np.round = half-even; astype(int)
  int dy = __double2int_rn(__dmul_rn(p.v, g.sy));
  dx = min(max(dx, 0), g.depth_w - 1);
  dy = min(max(dy, 0), g.depth_h - 1);
  return dy * g.depth_w + dx;
}

// IH:368-371
__device__ __forceinline__ bool depth_test(double z, uint16_t raw, const ImageGeom& g) {
  const double d = __dmul_rn(static_cast<double>(raw), g.depth_scale);
  return (z > 0.0) & (z < d);
}

template <bool STD, bool SINGLE = false>
__device__ __forceinline__ bool visible_exact(const double* __restrict__ w, const Intrinsics& K, const ImageGeom& g,
                                              const uint16_t* __restrict__ depth, double x, double y, double z,
                                              Proj& out) {
  out = project_exact<STD, SINGLE>(w, K, x, y, z);
  bool vis = false;
  if (in_image(out, g) && out.z > 0.0) {
    vis = depth_test(out.z, ldg_u16_nc(depth + depth_index(out, g)), g);
  }
  return vis;
}

// sample integers (visual_correspondence_qa_engine_coor_2_coor.py:341-344, depth_estimation_dot_engine.py:216-218,261)
__device__ __forceinline__ int permille(double coord, double extent) {
  return __double2int_rn(__dmul_rn(__ddiv_rn(coord, extent), 1000.0));
}
__device__ __forceinline__ int depth_mm(double z) { return __double2int_rn(__dmul_rn(z, 1000.0)); }

// ------------------------------------------------------------------------------------------------
// point loads
// ------------------------------------------------------------------------------------------------
template <int DT>
struct PtTraits;
template <>
struct PtTraits<0> {
  static constexpr int kBytes = 12;
  __device__ static __forceinline__ void load(const void* base, int64_t i, double& x, double& y, double& z) {
    const float* p = static_cast<const float*>(base) + 3 * i;
    x = p[0], y = p[1], z = p[2];
  }
};
template <>
struct PtTraits<1> {
  static constexpr int kBytes = 16;
  __device__ static __forceinline__ void load(const void* base, int64_t i, double& x, double& y, double& z) {
    const float4 q = static_cast<const float4*>(base)[i];
    x = q.x, y = q.y, z = q.z;
  }
};
template <>
struct PtTraits<2> {
  static constexpr int kBytes = 24;
  __device__ static __forceinline__ void load(const void* base, int64_t i, double& x, double& y, double& z) {
    const double* p = static_cast<const double*>(base) + 3 * i;
    x = p[0], y = p[1], z = p[2];
  }
};

// ------------------------------------------------------------------------------------------------
// 4x4 helpers for the pose prologue
// ------------------------------------------------------------------------------------------------
// C = A @ B with the same k-sequential FMA chain as the BLAS call behind IH:122.
__device__ inline void matmul4(const double* A, const double* B, double* C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double acc = __dmul_rn(A[i * 4 + 0], B[0 * 4 + j]);
      acc = __fma_rn(A[i * 4 + 1], B[1 * 4 + j], acc);
      acc = __fma_rn(A[i * 4 + 2], B[2 * 4 + j], acc);
      acc = __fma_rn(A[i * 4 + 3], B[3 * 4 + j], acc);
      C[i * 4 + j] = acc;
    }
}

// 4x4 inverse in EXACTLY the operation order of np.linalg.inv as built on OpenBLAS (dgesv = getrf + getrs):
//   * left-looking LU (getf2): column j first receives the earlier row swaps, then every entry has the dot product of
//     the already factored part subtracted — the dot is accumulated from 0 with FMAs in ascending k and subtracted once;
//     pivot = first largest |.| on or below the diagonal; the sub-diagonal is scaled by the RECIPROCAL of the pivot;
//   * forward substitution with fma(-l, y, s) in ascending k;
//   * back substitution with fma(-u, x, s) in DESCENDING k and a multiplication by the reciprocal of the diagonal.
// Found by exhaustive search over the plausible orders with exact-FMA emulation and verified bit-for-bit against
// np.linalg.inv on 1 860 matrices (ScanNet-like poses, TAPVid w2c, random affine and general matrices with pivoting,
// kilometre-scale translations): tests/test_oracle_golden.py::test_device_inverse_order_matches_numpy restates it.
// Returns false for a singular matrix.
__device__ inline bool inverse4(const double* M, double* inv) {
  double a[16];
  int perm[4] = {0, 1, 2, 3}, ipiv[4] = {0, 1, 2, 3};
  for (int i = 0; i < 16; ++i) a[i] = M[i];
  for (int j = 0; j < 4; ++j) {
    double b[4];
    for (int i = 0; i < 4; ++i) b[i] = a[i * 4 + j];
    for (int i = 0; i < j; ++i) {
      const int ip = ipiv[i];
      if (ip != i) {
        const double t = b[i];
        b[i] = b[ip];
        b[ip] = t;
      }
    }
    for (int i = 1; i < j; ++i) {
      double acc = 0.0;
      for (int k = 0; k < i; ++k) acc = __fma_rn(a[i * 4 + k], b[k], acc);
      b[i] = __dsub_rn(b[i], acc);
    }
    for (int r = j; r < 4; ++r) {
      double acc = 0.0;
      for (int k = 0; k < j; ++k) acc = __fma_rn(a[r * 4 + k], b[k], acc);
      b[r] = __dsub_rn(b[r], acc);
    }
    int jp = j;
    double best = fabs(b[j]);
    for (int r = j + 1; r < 4; ++r) {
      const double v = fabs(b[r]);
      if (v > best) best = v, jp = r;
    }
    ipiv[j] = jp;
    if (!(best > 0.0)) return false;
    if (jp != j) {
      const double t = b[j];
      b[j] = b[jp];
      b[jp] = t;
      for (int k = 0; k < j; ++k) {
        const double u = a[j * 4 + k];
        a[j * 4 + k] = a[jp * 4 + k];
        a[jp * 4 + k] = u;
      }
      const int tp = perm[j];
      perm[j] = perm[jp];
      perm[jp] = tp;
    }
    const double rp = __ddiv_rn(1.0, b[j]);
    for (int r = j + 1; r < 4; ++r) b[r] = __dmul_rn(b[r], rp);
    for (int i = 0; i < 4; ++i) a[i * 4 + j] = b[i];
  }
  for (int j = 0; j < 4; ++j) {  // solve L U x = P e_j
    double y[4], x[4];
    for (int i = 0; i < 4; ++i) {
      double s = (perm[i] == j) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s = __fma_rn(-a[i * 4 + k], y[k], s);
      y[i] = s;
    }
    for (int i = 3; i >= 0; --i) {
      double s = y[i];
      for (int k = 3; k > i; --k) s = __fma_rn(-a[i * 4 + k], x[k], s);
      x[i] = __dmul_rn(s, __ddiv_rn(1.0, a[i * 4 + i]));
    }
    for (int i = 0; i < 4; ++i) inv[i * 4 + j] = x[i];
  }
  return true;
}

}  // namespace msp
