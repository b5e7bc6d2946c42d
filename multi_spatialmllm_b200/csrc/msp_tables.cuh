// Visibility-table text on the device: the "values" column of the reference's canonical visibility parquet
// (make_visibility_info.py:59-65 — one row per frame with json.dumps(sorted visible point ids), one row per point with
// json.dumps(sorted ids of the frames that see it)) is produced straight from the bit masks, so the host never walks a
// Python list (the reference's 3-hour loop, MVI:103-118, and the json.dumps of ~10^7 integers per scene).
//
// Both tables are written in two phases: a size kernel (bytes per row), an exclusive scan by the caller (who needs the
// total to allocate the text anyway), and a write kernel that places every row at its offset.  Text format = Python's
// json.dumps defaults: "[" + ", ".join(items) + "]", integers in decimal, strings in double quotes (frame ids are
// restricted to characters json.dumps does not escape; the host checks that).  Byte-exact with the reference's rows.
#pragma once

#include <stdint.h>

namespace msp {

__device__ __forceinline__ int ndigits_u32(uint32_t v) {
  return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
       : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}

// bytes of ", ".join(str(base + b) for set bits b) + ", "  (every element is charged its separator; see the row formula)
__device__ __forceinline__ uint32_t json_word_bytes(uint32_t bits, uint32_t base) {
  if (bits == 0u) return 0u;
  const int dlo = ndigits_u32(base), dhi = ndigits_u32(base + 31u);
  if (dlo == dhi) return (uint32_t)__popc(bits) * (uint32_t)(dlo + 2);
  uint32_t t = 0;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1u;
    t += (uint32_t)ndigits_u32(base + (uint32_t)b) + 2u;
  }
  return t;
}

constexpr int kJsonThreads = 256;

// block-wide exclusive scan of one uint32 per thread (256 threads); returns the exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // s_warp may still be read by the previous tile
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kJsonThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) before += t;
    all += t;
  }
  total = all;
  return before + incl - v;
}

// ---- image_to_points: one CTA per requested frame row.
// Row text: '[' + for each visible point: digits + (", " | "]" for the last) ; "[]" when nothing is visible.
// Row bytes = sum over elements of (digits + 2) when non-empty (the '[' takes the place of the last element's missing
// separator byte and ']' of the other), 2 when empty.
template <bool WRITE>
__global__ void __launch_bounds__(kJsonThreads) k_json_point_lists(const uint32_t* __restrict__ mask, int64_t n,
                                                                   int64_t mask_stride, const int32_t* __restrict__ rows,
                                                                   const int64_t* __restrict__ offsets,
                                                                   int64_t* __restrict__ sizes, uint8_t* __restrict__ text) {
  __shared__ uint32_t s_warp[kJsonThreads / 32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int64_t f = rows ? (int64_t)rows[r] : (int64_t)r;
  const uint32_t* __restrict__ row = mask + f * mask_stride;
  const int nwords = (int)((n + 31) / 32);
  int64_t row_total = 0;
  uint8_t* out = nullptr;
  if (WRITE) {
    const int64_t o0 = offsets[r];
    row_total = offsets[r + 1] - o0;
    out = text + o0;
    if (tid == 0) {
      out[0] = '[';
      if (row_total == 2) out[1] = ']';
    }
  }
  int64_t base = 1;  // text position of the tile's first element
  for (int w0 = 0; w0 < nwords; w0 += kJsonThreads) {
    const int w = w0 + tid;
    uint32_t bits = (w < nwords) ? __ldg(row + w) : 0u;
    const int64_t p0 = (int64_t)w * 32;
    if (w < nwords && n - p0 < 32) bits &= (1u << (int)(n - p0)) - 1u;  // never trust pad bits
    const uint32_t mine = json_word_bytes(bits, (uint32_t)p0);
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan_256(mine, s_warp, tile_total);
    if (WRITE) {
      int64_t pos = base + excl;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        uint32_t v = (uint32_t)p0 + (uint32_t)b;
        const int d = ndigits_u32(v);
        for (int i = d - 1; i >= 0; --i) {
          out[pos + i] = (uint8_t)('0' + v % 10u);
          v /= 10u;
        }
        pos += d;
        if (pos + 1 == row_total) {
          out[pos] = ']';
        } else {
          out[pos] = ',';
          out[pos + 1] = ' ';
        }
        pos += 2;
      }
    }
    base += tile_total;
  }
  if (!WRITE && tid == 0) sizes[r] = (base == 1) ? 2 : base - 1;
}

// ---- point_to_images: one thread per point, frames walked in the caller's order (the reference sorts the id strings,
// MVI:116-118; the caller passes `rows` in that order).  Element text: '"' + id + '"' + (", " | "]").
// Row bytes = sum over elements of (len(id) + 4) when non-empty, 2 when empty.
template <bool WRITE>
__global__ void __launch_bounds__(kJsonThreads) k_json_image_lists(const uint32_t* __restrict__ mask, int64_t n,
                                                                   int64_t mask_stride, const int32_t* __restrict__ rows,
                                                                   int32_t R, const uint8_t* __restrict__ ids,
                                                                   int32_t id_pitch, const uint8_t* __restrict__ id_len,
                                                                   const int64_t* __restrict__ offsets,
                                                                   int64_t* __restrict__ sizes, uint8_t* __restrict__ text) {
  const int64_t p = (int64_t)blockIdx.x * kJsonThreads + threadIdx.x;
  if (p >= n) return;
  const int64_t word = p >> 5;  // uniform across the warp: one broadcast load per frame
  const uint32_t bit = (uint32_t)(p & 31);
  int64_t row_total = 0, pos = 1;
  uint8_t* out = nullptr;
  if (WRITE) {
    const int64_t o0 = offsets[p];
    row_total = offsets[p + 1] - o0;
    out = text + o0;
    out[0] = '[';
    if (row_total == 2) {
      out[1] = ']';
      return;  // nothing visible: no need to walk the frames again
    }
  }
  int64_t total = 0;
  constexpr int U = 8;  // independent mask loads in flight
  for (int r0 = 0; r0 < R; r0 += U) {
    uint32_t wv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int r = r0 + k;
      wv[k] = 0u;
      if (r < R) {
        const int64_t f = rows ? (int64_t)__ldg(rows + r) : (int64_t)r;
        wv[k] = __ldg(mask + f * mask_stride + word);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (!((wv[k] >> bit) & 1u)) continue;
      const int r = r0 + k;
      const int L = (int)__ldg(id_len + r);
      if (!WRITE) {
        total += L + 4;
      } else {
        const uint8_t* __restrict__ src = ids + (int64_t)r * id_pitch;
        out[pos++] = '"';
        for (int i = 0; i < L; ++i) out[pos + i] = __ldg(src + i);
        pos += L;
        out[pos++] = '"';
        if (pos + 1 == row_total) {
          out[pos] = ']';
        } else {
          out[pos] = ',';
          out[pos + 1] = ' ';
        }
        pos += 2;
      }
    }
  }
  if (!WRITE) sizes[p] = total ? total : 2;
}

}  // namespace msp
