// Vertex -> nearest pixel of the mesh-reprojection loss on the pixel LATTICE (find_nearest_neighbors, ind_BA,
// src/ops.py:60-71).
//
// The silhouette points of the reference are `where(seg > 0)` (src/trainer.py:443): integer (col, row) pairs in
// row-major order.  Such a set is a BITMAP, and a nearest-pixel query needs no point list at all: candidates are
// the set bits on square lattice rings around the vertex's own lattice point, their coordinates are implied by the
// bit position, |a|^2 is an exact integer, and the index `where` gave a pixel is its rank in the bitmap (row prefix +
// popcount).  Nothing is loaded per candidate: everything the walk reads sits in 21 KB of shared memory (the binned
// grid walk chased 16-byte point loads through L2).
//
// Every candidate is compared with the SAME fp32 expansion as the full scan (smaller index on equal values), a ring
// is skipped only when all of its pixels are provably farther than the best candidate by more than the rounding
// error of that expansion, and on a ring only the chord inside the circle of the best candidate is visited -- so the
// indices are those of the reference's full scan, bit for bit (tests/test_gpu_parity.py::
// test_mesh_grid_search_equals_brute_force, tests/test_gpu_baseline_sizes.py).  Images whose points are not such a
// list (k_lattice_build checks: integers in [0, LAT_N)^2, strictly increasing in row-major order) keep the binned-grid
// search of k_loss.cu.  (The other direction, pixel -> nearest vertex with the vertices binned to lattice cells, was
// built and measured too: 0.77 ms against 0.92 ms for the grid, but its per-image radix sort cost 0.32 ms against
// 0.19 ms for the grid build, and the row-major vertex order made THIS kernel slower; dropped.)
#include "mesh_common.cuh"

// Nearest set bit to `pos` in the 256-bit line `line` of table tb (-1: the line is empty).
__device__ __forceinline__ int nearest_in_line(const unsigned *tb, int line, int pos) {
  int bp = -1, bd = 1 << 30;
  for (int w = 0; w < LAT_W; ++w) {
    const unsigned bits = tb[line * LAT_W + w];
    if (!bits) continue;
    const int base = 32 * w;
    int pl = -1, ph = -1;
    if (pos >= base + 32) pl = base + 31 - __clz((int)bits);
    else if (pos < base) ph = base + __ffs((int)bits) - 1;
    else {
      const unsigned ml = bits & (~0u >> (31 - (pos - base))), mh = bits & (~0u << (pos - base));
      if (ml) pl = base + 31 - __clz((int)ml);
      if (mh) ph = base + __ffs((int)mh) - 1;
    }
    if (pl >= 0 && pos - pl < bd) {
      bd = pos - pl;
      bp = pl;
    }
    if (ph >= 0 && ph - pos < bd) {
      bd = ph - pos;
      bp = ph;
    }
  }
  return bp;
}

// rounding-error bound of the expansion (magnitudes up to max(|a|^2, |b|^2)), as in the grid search of k_loss.cu
__device__ __forceinline__ float lattice_margin(const float *gparam, int i) {
  const float *gpB = gparam + ((size_t)i * 2 + 0) * GP_STRIDE, *gpA = gparam + ((size_t)i * 2 + 1) * GP_STRIDE;
  return 32.0f * 5.9604645e-8f * fmaxf(fmaxf(gpA[6], gpB[6]), 1.0f);
}

// ------------------------------------------------------------------- vertex -> nearest pixel
#define LT 512   // threads per CTA (the tables are loaded once per CTA)
__global__ void __launch_bounds__(LT) k_mesh_ba_lat(int V, const int *__restrict__ offsets, float *__restrict__ vdist,
                                                    float *__restrict__ d_sil, int *__restrict__ ind_ba,
                                                    const float *__restrict__ gparam, const float4 *__restrict__ sortedV,
                                                    const unsigned char *__restrict__ lat, const int *__restrict__ lat_ok) {
  __shared__ __align__(16) unsigned char s_lat[LAT_BYTES];
  const int i = blockIdx.y;
  if (!lat_ok[i]) return;       // this image's points are not a row-major pixel list: k_mesh_ba<true> handles it
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(lat + (size_t)i * LAT_BYTES);
    for (int k = threadIdx.x; k < LAT_BYTES / 16; k += LT) reinterpret_cast<uint4 *>(s_lat)[k] = src[k];
    __syncthreads();
  }
  const unsigned *bm = reinterpret_cast<const unsigned *>(s_lat), *bmT = bm + LAT_BM_WORDS;
  const int *pref = reinterpret_cast<const int *>(bmT + LAT_BM_WORDS);
  const unsigned char *cd = reinterpret_cast<const unsigned char *>(pref + LAT_N);
  const int np = offsets[i + 1] - offsets[i];
  const int slot = blockIdx.x * LT + threadIdx.x;
  if (slot >= V) return;
  // binned vertex order: the lanes of a warp query neighbouring positions
  const float4 q = sortedV[(size_t)i * V + slot];
  const int b = __float_as_int(q.w);
  const float bx = q.x, by = q.y, b2 = q.z;
  float best = 3.4e38f;
  int ai = -1, apx = 0, apy = 0;
  if (np > 0) {
    const float margin = lattice_margin(gparam, i);
    const float qcx = fminf(fmaxf(bx, 0.0f), (float)(LAT_N - 1)), qcy = fminf(fmaxf(by, 0.0f), (float)(LAT_N - 1));
    const float outside2 = (bx - qcx) * (bx - qcx) + (by - qcy) * (by - qcy);
    const int gx = (int)rintf(qcx), gy = (int)rintf(qcy);
    auto cand = [&](int px, int py) {
      const float ax = (float)px, ay = (float)py;
      const float a2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
      const float d = d2_expand(ax, ay, a2, bx, by, b2);
      if (d <= best) {
        // rank of the pixel in the row-major list
        int id = pref[py];
        const int w = px >> 5;
        for (int qq = 0; qq < w; ++qq) id += __popc(bm[py * LAT_W + qq]);
        id += __popc(bm[py * LAT_W + w] & ((1u << (px & 31)) - 1u));
        if (d < best || id < ai) {
          best = d;
          ai = id;
          apx = px;
          apy = py;
        }
      }
    };
    // set bits lo..hi of the 256-bit line `line` of table tb; horizontal: line = row, bit = column
    auto scan_line = [&](const unsigned *tb, int line, int lo, int hi, bool horizontal) {
      if (best < 1e30f) {
        // only the chord of the line inside the circle of the best candidate so far (+ the rounding margin) can matter
        const float off = (float)line - (horizontal ? by : bx), mid = horizontal ? bx : by;
        const float rem = best + margin - off * off;
        if (rem < 0.0f) return;
        const float w = sqrtf(rem) * 1.0001f + 1e-3f;
        lo = max(lo, (int)ceilf(mid - w));
        hi = min(hi, (int)floorf(mid + w));
      }
      if (lo > hi) return;
      const int wa = lo >> 5, wb = hi >> 5;
      for (int w = wa; w <= wb; ++w) {
        unsigned bits = tb[line * LAT_W + w];
        if (w == wa) bits &= ~0u << (lo & 31);
        if (w == wb) bits &= ~0u >> (31 - (hi & 31));
        while (bits) {
          const int p = 32 * w + __ffs((int)bits) - 1;
          bits &= bits - 1;
          if (horizontal) cand(p, line);
          else cand(line, p);
        }
      }
    };
    // every pixel lies in a coarse cell at Chebyshev distance >= c from the query's cell, i.e. on a lattice ring
    // >= 4 (c - 1) + 1
    const int c0 = cd[(gy >> 2) * LAT_C + (gx >> 2)];
    // (a cell c0 cells away along some axis starts 4 c0 - o lattice points away, o = g's offset inside its own cell
    // towards it: at least 4 c0 - 3, and 4 c0 - max(o_x, 3 - o_x, o_y, 3 - o_y) whichever the direction)
    int k = c0 >= 1 ? 4 * c0 - max(max(gx & 3, 3 - (gx & 3)), max(gy & 3, 3 - (gy & 3))) : 0;
    const int kmax = max(max(gx, LAT_N - 1 - gx), max(gy, LAT_N - 1 - gy));
    if (k == 0) {
      // Rings 0 and 1 together, without loops (most queries lie inside the silhouette and end here): the 3 x 3 block of
      // lattice points around g as three 3-bit fields, compared in row-major order with a strict '<' (the first of
      // equal distances has the smallest index); the rank is computed once, for the winner.
      float bd = 3.4e38f;
      int bpx = -1, bpy = 0;
      const int xs = max(gx - 1, 0), w0 = xs >> 5;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int py = gy + dy;
        if (py < 0 || py >= LAT_N) continue;
        const unsigned lo = bm[py * LAT_W + w0], hi = w0 + 1 < LAT_W ? bm[py * LAT_W + w0 + 1] : 0u;
        const unsigned bits = __funnelshift_r(lo, hi, xs & 31) & (gx == 0 ? 3u : 7u);   // columns xs, xs + 1, xs + 2
        const float ay = (float)py;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float ax = (float)(xs + dx);
          const float d = d2_expand(ax, ay, __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), bx, by, b2);
          if ((bits >> dx & 1u) && d < bd) {
            bd = d;
            bpx = xs + dx;
            bpy = py;
          }
        }
      }
      if (bpx >= 0) cand(bpx, bpy);
      k = 2;
    } else if (c0 >= 2) {
      // A far query: before the walk, take the nearest pixel of the query's row and of its column (clamped to the
      // bounding box of the pixels) as first candidates, so that the chord restriction applies from the first ring on.
      const float *gpA = gparam + ((size_t)i * 2 + 1) * GP_STRIDE;
      const int ry = min(max(gy, (int)gpA[1]), (int)gpA[3]), rx = min(max(gx, (int)gpA[0]), (int)gpA[2]);
      const int sx = nearest_in_line(bm, ry, gx);
      if (sx >= 0) cand(sx, ry);
      const int sy = nearest_in_line(bmT, rx, gy);
      if (sy >= 0) cand(rx, sy);
    }
    for (; k <= kmax; ++k) {
      // unvisited pixels are on rings >= k: |p - g|_inf >= k and |qc - g|_inf <= 0.5, so |qc - p| >= k - 0.5, and
      // |q - p|^2 >= |q - qc|^2 + |qc - p|^2 (qc is the projection of q onto the lattice's box)
      const float lb = (float)k - 0.5f;
      if (lb * lb * 0.9999f + outside2 > best + margin) break;
      const int xa = max(gx - k, 0), xb = min(gx + k, LAT_N - 1);
      if (gy - k >= 0) scan_line(bm, gy - k, xa, xb, true);
      if (gy + k < LAT_N) scan_line(bm, gy + k, xa, xb, true);
      const int ya = max(gy - k + 1, 0), yb = min(gy + k - 1, LAT_N - 1);
      if (gx - k >= 0) scan_line(bmT, gx - k, ya, yb, false);
      if (gx + k < LAT_N) scan_line(bmT, gx + k, ya, yb, false);
    }
  }
  float dist = 0.f, gxo = 0.f, gyo = 0.f;
  if (ind_ba) ind_ba[(size_t)i * V + b] = ai;
  if (ai >= 0) {
    const float dx = bx - (float)apx, dy = by - (float)apy;
    dist = sqrtf(dx * dx + dy * dy);
    if (dist > 0.f) {   // tf.norm's gradient is NaN at exactly 0 (measure zero); 0 here
      gxo = dx / dist;
      gyo = dy / dist;
    }
  }
  if (d_sil) {
    d_sil[((size_t)i * V + b) * 2 + 0] = gxo;
    d_sil[((size_t)i * V + b) * 2 + 1] = gyo;
  }
  vdist[(size_t)i * V + b] = dist;
}

// ------------------------------------------------------------------------------------- host
size_t mesh_lattice_workspace(int B) { return (size_t)B * LAT_BYTES + (size_t)B * 4 + 256; }

// sortedB: the vertices in the binned order of k_grid_build (neighbouring lanes query neighbouring positions)
int launch_mesh_lattice_search(smplb_ctx *c, int B, int V, const int *offsets, const void *ws, const float *gparam,
                               const float4 *sortedB, float *vdist, float *d_sil, int *ind_ba) {
  const unsigned char *lat = (const unsigned char *)ws;
  const int *lat_ok = (const int *)(lat + (size_t)B * LAT_BYTES);
  LAUNCH(c, "mesh_nn_vertex_to_pixel_lattice", dim3(cdiv(V, LT), B), LT, 0, k_mesh_ba_lat, V, offsets, vdist, d_sil, ind_ba, gparam,
         sortedB, lat, lat_ok);
  return 0;
}
