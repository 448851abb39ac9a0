// Loss kernels: keypoint loss, mesh-reprojection (bidirectional nearest-neighbour) loss and
// the WGAN-GP gradient-penalty reduction.  All reductions are fixed-order (deterministic).
//
//   kp_reprojection_loss      src/ops.py:35-47
//   find_nearest_neighbors    src/ops.py:60-71
//   bidirectional_dist        src/ops.py:83-102
//   mesh_reprojection_loss    src/ops.py:117-137
//   compute_gradient_penalty  src/ops.py:153-172
#include "smplb_internal.h"
#include "mesh_common.cuh"

// ---------------------------------------------------------------------------- keypoint loss
// Stand-alone kp loss on caller-supplied predictions: per-body partial sums.
__global__ void k_kp_loss(int B, int K, const float *__restrict__ kp_gt, const float *__restrict__ kp_pred,
                          float *__restrict__ dkp, float *__restrict__ part, int *__restrict__ cnt) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float tl = 0.f;
  int tc = 0;
  for (int k = 0; k < K; ++k) {
    size_t bk = (size_t)b * K + k;
    float gx = kp_gt[bk * 3 + 0], gy = kp_gt[bk * 3 + 1], vis = kp_gt[bk * 3 + 2];
    float dx = kp_pred[bk * 2 + 0] - gx, dy = kp_pred[bk * 2 + 1] - gy;
    tl += vis * fabsf(dx) + vis * fabsf(dy);
    tc += (vis != 0.0f) ? 2 : 0;
    if (dkp) {
      dkp[bk * 2 + 0] = vis * (float)((dx > 0.f) - (dx < 0.f));
      dkp[bk * 2 + 1] = vis * (float)((dy > 0.f) - (dy < 0.f));
    }
  }
  part[b] = tl;
  cnt[b] = tc;
}

// Sum of the per-body partials: one block, fixed order.  256 threads (8 K registers), so the
// block can be placed on an SM that a persistent tensor-core CTA occupies: a 1024-thread block
// could not, and while it waited at the head of the queue it held back every later launch of
// every stream (tools/timeline.py showed the whole keypoint path stalling ~95 us on it).
#define RK_THREADS 256
__device__ __forceinline__ void reduce_kp_block(int B, const float *__restrict__ part, const int *__restrict__ cnt,
                                                float *red, long long *redc) {
  const int t = threadIdx.x;
  float s = 0.f;
  long long c = 0;
  for (int i = t; i < B; i += RK_THREADS) {
    s += part[i];
    c += cnt[i];
  }
  red[t] = s;
  redc[t] = c;
  __syncthreads();
  for (int o = RK_THREADS / 2; o > 0; o >>= 1) {
    if (t < o) {
      red[t] += red[t + o];
      redc[t] += redc[t + o];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RK_THREADS) k_reduce_kp(int B, const float *__restrict__ part,
                                                          const int *__restrict__ cnt, float *__restrict__ abs_sum,
                                                          long long *__restrict__ num_present,
                                                          float *__restrict__ cnt_as_float) {
  __shared__ float red[RK_THREADS];
  __shared__ long long redc[RK_THREADS];
  reduce_kp_block(B, part, cnt, red, redc);
  if (threadIdx.x == 0) {
    *abs_sum = red[0];
    *num_present = redc[0];
    if (cnt_as_float) *cnt_as_float = (float)redc[0];   // exact below 2^24; the all-reduced form
  }
}

// Single-GPU, keypoint-loss-only steps: reduce and finalize in one launch (no all-reduce or
// mesh term has to come in between).
__global__ void __launch_bounds__(RK_THREADS) k_reduce_finalize(int B, const float *__restrict__ part,
                                                                const int *__restrict__ cnt, float w_kp,
                                                                long long count_override, float *__restrict__ scal,
                                                                long long *__restrict__ kp_cnt, float *__restrict__ out) {
  __shared__ float red[RK_THREADS];
  __shared__ long long redc[RK_THREADS];
  reduce_kp_block(B, part, cnt, red, redc);
  if (threadIdx.x == 0) {
    long long den = count_override > 0 ? count_override : redc[0];
    float kp = den > 0 ? red[0] / (float)den : 0.0f;
    scal[0] = red[0];
    scal[1] = (float)redc[0];
    out[0] = red[0];
    out[1] = (float)redc[0];
    out[2] = 0.0f;
    out[3] = w_kp * kp;
    *kp_cnt = den;
  }
}

// scal = {kp abs_sum, kp num_present as float, mesh sum} -- already all-reduced over the ranks
// when a communicator is attached.  loss_parts = {abs_sum, num_present, mesh sum,
// w_kp * abs_sum / num_present + w_mesh * mesh}; kp_cnt receives the count the backward
// divides by (count_override > 0 wins: single-process emulation of shards).
__global__ void k_finalize_loss(float w_kp, float w_mesh, long long count_override, int have_mesh,
                                const float *__restrict__ scal, long long *__restrict__ kp_cnt,
                                float *__restrict__ out) {
  long long den = count_override > 0 ? count_override : (long long)scal[1];
  float kp = den > 0 ? scal[0] / (float)den : 0.0f;
  float ml = have_mesh ? scal[2] : 0.0f;
  out[0] = scal[0];
  out[1] = scal[1];
  out[2] = ml;
  out[3] = w_kp * kp + w_mesh * ml;
  *kp_cnt = den;
}

// ------------------------------------------------------------------------------- mesh loss
#define MT 256     // threads per CTA of the NN kernels
#define MTILE 1024 // points staged per shared-memory tile

// ---- uniform-grid acceleration of the nearest-neighbour search -----------------------------
// The reference scans every (pixel, vertex) pair.  Here each point set of each image is binned
// into a GRID_G x GRID_G grid over its bounding box and a query walks Chebyshev rings of cells
// around its own cell, stopping once every unvisited cell is provably farther than the best
// candidate.  Candidates are compared with the SAME fp32 expansion as the brute-force scan
// (and the smaller index on equal values), and the stopping rule carries a margin that bounds
// the rounding error of that expansion, so the result is bit-identical to the full scan
// (tests/test_gpu_parity.py::test_mesh_grid_search_equals_brute_force) at a few percent of the pairs.
// 64 x 64 cells (~2 points per cell for 6890 vertices, ~6 for a 10 k-pixel silhouette): 32 x 32 made
// the pixel -> vertex search 2.9x slower, 96 x 96 the vertex -> pixel search 1.4x slower.
#define GRID_G 64
#define GRID_NC (GRID_G * GRID_G)
#define GP_STRIDE 8   // x0, y0, x1, y1, inv_h, h, max |p|^2, -
// Most of a far query's ring walk crosses EMPTY cells (a vertex 50 px outside the silhouette walks ~17 rings,
// ~600 cell ranges, before it meets the first pixel).  Per (image, set) the build kernel therefore also writes
//   occ [y]  : bit x set if cell (x, y) holds a point          (64 x 64 cells = one 64-bit word per row)
//   occT[x]  : the same by column
//   cdist[c] : Chebyshev distance in cells from cell c to the nearest non-empty cell
// and the walk (a) starts at ring cdist[c] -- every ring before it is empty -- and (b) visits only the
// non-empty cells of a ring: two row masks and two column masks per ring.  The set of points compared
// is unchanged, so the result stays bit-identical to the full scan.
#define GRID_AUX_BYTES (2 * GRID_G * 8 + GRID_NC)   // occ, occT, cdist

// One CTA per (set, image): set 0 = projected vertices, set 1 = silhouette pixels (the two CTAs of an image run side by
// side: building the pixel tables in a launch of its own cost as much as this one).
__global__ void __launch_bounds__(256) k_grid_build(int V, const float *__restrict__ pts, const int *__restrict__ offsets,
                                                    const float *__restrict__ sil_pred, float *__restrict__ gparam,
                                                    int *__restrict__ gstart, float4 *__restrict__ sortedB,
                                                    float4 *__restrict__ sortedA, unsigned char *__restrict__ gaux,
                                                    unsigned char *__restrict__ lat, int *__restrict__ lat_ok) {
  // One buffer, two views (a CTA bins points OR builds the lattice tables; a pixel set that is not a pixel list does the
  // second, then the first): 22.5 KB instead of 40 KB lets 8 CTAs share an SM, i.e. the 2 B CTAs run in fewer rounds.
  __shared__ __align__(16) unsigned char s_raw[GRID_NC * 4 + 4 * 256 * 4 + 256 * 4 + 2 * GRID_G * 8];
  int *hist = reinterpret_cast<int *>(s_raw);                                           // [GRID_NC]       | bm, bmT
  float(*red)[256] = reinterpret_cast<float(*)[256]>(s_raw + GRID_NC * 4);              // [4][256]
  int *scan = reinterpret_cast<int *>(s_raw + GRID_NC * 4 + 4096);                      // [256]
  unsigned long long *s_occ = reinterpret_cast<unsigned long long *>(s_raw + GRID_NC * 4 + 4096 + 1024);   // [GRID_G]
  unsigned long long *s_tmp = s_occ + GRID_G;                                           // [GRID_G]
  static_assert(2 * LAT_BM_WORDS * 4 <= GRID_NC * 4 && LAT_C == GRID_G, "the lattice tables fit the histogram's space");
  int sel = blockIdx.x, i = blockIdx.y, t = threadIdx.x;
  // the pixel set: tables for the lattice search first (k_mesh_lattice.cu); if the points are a row-major pixel list
  // that is all this image's pixels need, otherwise they are binned like the vertices
  if (sel == 1 && lat &&
      lattice_build_image(i, pts, offsets, lat, lat_ok, gparam, reinterpret_cast<unsigned *>(s_raw),
                          reinterpret_cast<unsigned *>(s_raw) + LAT_BM_WORDS, scan, s_occ, s_tmp))
    return;
  int p0 = offsets[i];
  int n = sel == 0 ? V : offsets[i + 1] - p0;
  const float *src = sel == 0 ? sil_pred + (size_t)i * V * 2 : pts + (size_t)p0 * 2;
  float4 *dst = sel == 0 ? sortedB + (size_t)i * V : sortedA + p0;
  float *gp = gparam + ((size_t)i * 2 + sel) * GP_STRIDE;
  int *gs = gstart + ((size_t)i * 2 + sel) * (GRID_NC + 1);
  float mnx = 3.4e38f, mny = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f;
  // (points are read as 8-byte pairs, four iterations' loads in flight: these loops were bound by one load latency each)
  const float2 *src2 = reinterpret_cast<const float2 *>(src);
#pragma unroll 4
  for (int k = t; k < n; k += 256) {
    const float2 pt = src2[k];
    float x = pt.x, y = pt.y;
    mnx = fminf(mnx, x);
    mny = fminf(mny, y);
    mxx = fmaxf(mxx, x);
    mxy = fmaxf(mxy, y);
  }
  red[0][t] = mnx;
  red[1][t] = mny;
  red[2][t] = mxx;
  red[3][t] = mxy;
  for (int k = t; k < GRID_NC; k += 256) hist[k] = 0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) {
      red[0][t] = fminf(red[0][t], red[0][t + o]);
      red[1][t] = fminf(red[1][t], red[1][t + o]);
      red[2][t] = fmaxf(red[2][t], red[2][t + o]);
      red[3][t] = fmaxf(red[3][t], red[3][t + o]);
    }
    __syncthreads();
  }
  float x0 = red[0][0], y0 = red[1][0], x1 = red[2][0], y1 = red[3][0];
  if (n == 0) {
    x0 = y0 = x1 = y1 = 0.f;
  }
  float ext = fmaxf(x1 - x0, y1 - y0);
  float h = ext > 0.f ? ext / GRID_G : 1.0f;
  float inv_h = 1.0f / h;
#pragma unroll 4
  for (int k = t; k < n; k += 256) {
    const float2 pt = src2[k];
    float x = pt.x, y = pt.y;
    int cx = min(GRID_G - 1, max(0, (int)((x - x0) * inv_h)));
    int cy = min(GRID_G - 1, max(0, (int)((y - y0) * inv_h)));
    atomicAdd(&hist[cy * GRID_G + cx], 1);
  }
  __syncthreads();
  // occupancy masks and the cell distance transform (see GRID_AUX_BYTES)
  {
    unsigned char *aux = gaux + ((size_t)i * 2 + sel) * GRID_AUX_BYTES;
    unsigned long long *occ = reinterpret_cast<unsigned long long *>(aux), *occT = occ + GRID_G;
    unsigned char *cdist = aux + 2 * GRID_G * 8;
    if (t < GRID_G) {
      unsigned long long m = 0;
      for (int x = 0; x < GRID_G; ++x) m |= (unsigned long long)(hist[t * GRID_G + x] > 0) << x;
      s_occ[t] = m;
      occ[t] = m;
    } else if (t < 2 * GRID_G) {
      const int x = t - GRID_G;
      unsigned long long m = 0;
      for (int y = 0; y < GRID_G; ++y) m |= (unsigned long long)(hist[y * GRID_G + x] > 0) << y;
      occT[x] = m;
    }
    __syncthreads();
    static_assert(GRID_G == 64, "chebyshev_cells64 is written for 64 x 64 cells");
    chebyshev_cells64(s_occ, s_tmp, cdist);
  }
  // exclusive scan of the cell counts: GRID_NC / 256 cells per thread + block scan
  constexpr int CPT = GRID_NC / 256;
  int cs = 0;
  for (int q = 0; q < CPT; ++q) cs += hist[CPT * t + q];
  scan[t] = cs;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int v = t >= o ? scan[t - o] : 0;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  int base = scan[t] - cs;
  __syncthreads();
  for (int q = 0; q < CPT; ++q) {     // hist now holds the running cursors
    int cq = hist[CPT * t + q];
    hist[CPT * t + q] = base;
    gs[CPT * t + q] = base;
    base += cq;
  }
  if (t == 255) gs[GRID_NC] = n;
  __syncthreads();
#pragma unroll 4
  for (int k = t; k < n; k += 256) {
    const float2 pt = src2[k];
    float x = pt.x, y = pt.y;
    int cx = min(GRID_G - 1, max(0, (int)((x - x0) * inv_h)));
    int cy = min(GRID_G - 1, max(0, (int)((y - y0) * inv_h)));
    int pos = atomicAdd(&hist[cy * GRID_G + cx], 1);
    dst[pos] = make_float4(x, y, __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __int_as_float(k));
  }
  if (t == 0) {
    gp[0] = x0;
    gp[1] = y0;
    gp[2] = x1;
    gp[3] = y1;
    gp[4] = inv_h;
    gp[5] = h;
    gp[6] = fmaxf(x0 * x0, x1 * x1) + fmaxf(y0 * y0, y1 * y1);
    gp[7] = 0.f;
  }
}

// Nearest candidate of query q in a binned set.  QUERY_IS_A: q is a pixel (a), candidates are
// vertices (b); otherwise the roles in the expansion are swapped.  Returns the candidate's
// ORIGINAL index (smaller index on equal values == first index of a full in-order scan).
template <bool QUERY_IS_A>
__device__ __forceinline__ void grid_search(float qx, float qy, float q2, const float4 *__restrict__ sorted,
                                            const int *__restrict__ gs, const float *__restrict__ gp, float margin,
                                            const unsigned long long *s_occ, const unsigned long long *s_occT,
                                            const unsigned char *s_cdist, float &best, int &bi) {
  float x0 = gp[0], y0 = gp[1], x1 = gp[2], y1 = gp[3], inv_h = gp[4], h = gp[5];
  float qcx = fminf(fmaxf(qx, x0), x1), qcy = fminf(fmaxf(qy, y0), y1);   // projection onto the bounding box
  float outside2 = (qx - qcx) * (qx - qcx) + (qy - qcy) * (qy - qcy);
  int cx = min(GRID_G - 1, max(0, (int)((qcx - x0) * inv_h)));
  int cy = min(GRID_G - 1, max(0, (int)((qcy - y0) * inv_h)));
  best = 3.4e38f;
  bi = 0x7fffffff;
  auto scan = [&](int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
      float4 cnd = sorted[k];
      float d = QUERY_IS_A ? d2_expand(qx, qy, q2, cnd.x, cnd.y, cnd.z) : d2_expand(cnd.x, cnd.y, cnd.z, qx, qy, q2);
      int id = __float_as_int(cnd.w);
      if (d < best || (d == best && id < bi)) {
        best = d;
        bi = id;
      }
    }
  };
  // the non-empty cells of row y between columns xa..xb, as runs of consecutive cells (contiguous point ranges)
  auto scan_row = [&](int y, int xa, int xb) {
    unsigned long long m = s_occ[y] & ((~0ull >> (63 - xb)) & (~0ull << xa));
    while (m) {
      const int xs = __ffsll((long long)m) - 1;
      const unsigned long long run = ~(m >> xs);                          // first zero above xs ends the run
      const int len = run ? __ffsll((long long)run) - 1 : 64 - xs;
      scan(gs[y * GRID_G + xs], gs[y * GRID_G + xs + len]);
      m = (xs + len >= 64) ? 0ull : (m & (~0ull << (xs + len)));
    }
  };
  // the non-empty cells of column x between rows ya..yb
  auto scan_col = [&](int x, int ya, int yb) {
    if (ya > yb) return;
    unsigned long long m = s_occT[x] & ((~0ull >> (63 - yb)) & (~0ull << ya));
    while (m) {
      const int y = __ffsll((long long)m) - 1;
      m &= m - 1;
      scan(gs[y * GRID_G + x], gs[y * GRID_G + x + 1]);
    }
  };
  const int rmax = max(max(cx, GRID_G - 1 - cx), max(cy, GRID_G - 1 - cy));
  const int r0 = s_cdist[cy * GRID_G + cx];      // every ring before r0 is empty
  if (r0 > rmax) return;                          // an empty set
  int r = r0;
  if (r0 <= 1) {
    // Rings 0 and 1 together (ring 0 alone can never satisfy the stopping rule, 0 * h): the query's own cell first,
    // then only those of its eight neighbours whose RECTANGLE can hold a point closer than the best so far.  With
    // ~12 vertices per cell and the nearest one ~1 px away, a pixel compares ~30 candidates instead of the ~100 of the
    // whole 3 x 3 block.  The cells of a row are one contiguous range; slack covers the rounding of the cell assignment.
    const int c00 = gs[cy * GRID_G + cx], c01 = gs[cy * GRID_G + cx + 1];
    int lo[3], hi[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // range bounds of the three rows, loaded before the first scan
      const int y = min(max(cy - 1 + j, 0), GRID_G - 1);
      lo[j] = gs[y * GRID_G + max(cx - 1, 0)];
      hi[j] = gs[y * GRID_G + min(cx + 1, GRID_G - 1) + 1];
    }
    scan(c00, c01);
    const float slack = 1e-4f * h + 1e-6f * (fabsf(x0) + fabsf(x1) + fabsf(y0) + fabsf(y1));
    const float fx = qcx - (x0 + (float)cx * h), fy = qcy - (y0 + (float)cy * h);   // position inside the own cell
    const float dl = fmaxf(fx - slack, 0.0f), dr = fmaxf(h - fx - slack, 0.0f);     // distance to the left / right neighbours
    const float du = fmaxf(fy - slack, 0.0f), dd = fmaxf(h - fy - slack, 0.0f);     // ... to the rows above / below
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int y = cy - 1 + j;
      if (y < 0 || y >= GRID_G) continue;
      const float ry = j == 0 ? du : (j == 2 ? dd : 0.0f);
      const float lim = best + margin - outside2;
      const bool mid_ok = j != 1 && ry * ry * 0.9999f <= lim;
      const bool left_ok = cx > 0 && (dl * dl + ry * ry) * 0.9999f <= lim;
      const bool right_ok = cx < GRID_G - 1 && (dr * dr + ry * ry) * 0.9999f <= lim;
      if (j == 1) {
        if (left_ok) scan(lo[1], c00);
        if (right_ok) scan(c01, hi[1]);
      } else if (mid_ok) {
        // (left / right can only pass if the middle cell does)
        const int a = left_ok ? lo[j] : gs[y * GRID_G + cx], b = right_ok ? hi[j] : gs[y * GRID_G + cx + 1];
        scan(a, b);
      }
    }
    r = 1;
  } else {
    --r;   // the loop below visits ring r + 1 first
  }
  for (;;) {
    // every unvisited cell is at Chebyshev ring >= r + 1: its points are >= r * h from the
    // projected query (0.9999 covers the rounding of the cell assignment)
    const float rh = (float)r * h;
    if (rh * rh * 0.9999f + outside2 > best + margin) break;
    if (++r > rmax) break;
    const int xa = max(cx - r, 0), xb = min(cx + r, GRID_G - 1);
    if (cy - r >= 0) scan_row(cy - r, xa, xb);
    if (cy + r < GRID_G) scan_row(cy + r, xa, xb);
    const int ya = max(cy - r + 1, 0), yb = min(cy + r - 1, GRID_G - 1);
    if (cx - r >= 0) scan_col(cx - r, ya, yb);
    if (cx + r < GRID_G) scan_col(cx + r, ya, yb);
  }
}

// pixel -> nearest vertex (ind_AB, ops.py:68): thread = pixel a of image i, scans all vertices
// in index order with a strict '<' so the first minimal index wins (tf.argmin).  Adds the L1
// term |A_a - B_nn| (ops.py:98) to a per-CTA partial and the integer sign sums of its gradient
// to cnt[i][nn][0..1].
template <bool GRID>
__global__ void __launch_bounds__(MT) k_mesh_ab(int V, const float *__restrict__ pts, const int *__restrict__ offsets,
                                                const float *__restrict__ sil_pred, float *__restrict__ part,
                                                int *__restrict__ cnt, int *__restrict__ ind_ab,
                                                const float *__restrict__ gparam, const int *__restrict__ gstart,
                                                const float4 *__restrict__ sortedB, const unsigned char *__restrict__ gaux) {
  __shared__ float4 s4[GRID ? 1 : MTILE];   // (x, y, |.|^2, -) of the staged vertices: one broadcast LDS.128 per pair
  __shared__ float red[MT];
  __shared__ __align__(16) unsigned char s_aux[GRID ? GRID_AUX_BYTES : 16];
  int i = blockIdx.y;
  if (GRID) {   // occupancy masks + cell distances of this image's VERTEX grid
    const uint4 *src = reinterpret_cast<const uint4 *>(gaux + ((size_t)i * 2 + 0) * GRID_AUX_BYTES);
    for (int k = threadIdx.x; k < GRID_AUX_BYTES / 16; k += MT) reinterpret_cast<uint4 *>(s_aux)[k] = src[k];
    __syncthreads();
  }
  int p0 = offsets[i], np = offsets[i + 1] - p0;
  const float *Bv = sil_pred + (size_t)i * V * 2;
  float total = 0.f;
  int nchunks = (np + MT - 1) / MT;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    int a = ch * MT + threadIdx.x;
    bool ok = a < np;
    float ax = 0.f, ay = 0.f;
    if (ok) {
      ax = pts[(size_t)(p0 + a) * 2 + 0];
      ay = pts[(size_t)(p0 + a) * 2 + 1];
    }
    float a2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
    float best = 3.4e38f;
    int bi = 0;
    if (GRID) {
      const float *gpB = gparam + ((size_t)i * 2 + 0) * GP_STRIDE, *gpA = gparam + ((size_t)i * 2 + 1) * GP_STRIDE;
      float margin = 32.0f * 5.9604645e-8f * fmaxf(gpA[6], gpB[6]);
      if (ok)
        grid_search<true>(ax, ay, a2, sortedB + (size_t)i * V, gstart + ((size_t)i * 2 + 0) * (GRID_NC + 1), gpB, margin,
                          reinterpret_cast<const unsigned long long *>(s_aux),
                          reinterpret_cast<const unsigned long long *>(s_aux) + GRID_G, s_aux + 2 * GRID_G * 8, best, bi);
    } else
    for (int v0 = 0; v0 < V; v0 += MTILE) {
      int nv = min(MTILE, V - v0);
      __syncthreads();
      for (int q = threadIdx.x; q < nv; q += MT) {
        float bx = Bv[(size_t)(v0 + q) * 2 + 0], by = Bv[(size_t)(v0 + q) * 2 + 1];
        s4[q] = make_float4(bx, by, __fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), 0.f);
      }
      __syncthreads();
#pragma unroll 4
      for (int q = 0; q < nv; ++q) {
        float4 bq = s4[q];
        float d = d2_expand(ax, ay, a2, bq.x, bq.y, bq.z);
        if (d < best) {
          best = d;
          bi = v0 + q;
        }
      }
    }
    if (ok) {
      if (ind_ab) ind_ab[p0 + a] = bi;
      float bx = Bv[(size_t)bi * 2 + 0], by = Bv[(size_t)bi * 2 + 1];
      float dx = ax - bx, dy = ay - by;
      total += fabsf(dx) + fabsf(dy);
      if (cnt) {
        // d |A_a - B_nn| / d B_nn = -sign(A_a - B_nn): an integer, accumulated exactly
        int sxg = (dx < 0.f) - (dx > 0.f), syg = (dy < 0.f) - (dy > 0.f);
        if (sxg) atomicAdd(&cnt[((size_t)i * V + bi) * 2 + 0], sxg);
        if (syg) atomicAdd(&cnt[((size_t)i * V + bi) * 2 + 1], syg);
      }
    }
  }
  float bs = block_sum(total, red);
  if (threadIdx.x == 0) part[(size_t)i * gridDim.x + blockIdx.x] = bs;
}

// vertex -> nearest pixel (ind_BA, ops.py:69): thread = vertex b, scans all pixels of image i
// in order.  Adds the L2 term ||B_b - A_nn|| (ops.py:92) and writes its gradient (unit vector).
template <bool GRID>
__global__ void __launch_bounds__(MT) k_mesh_ba(int V, const float *__restrict__ pts, const int *__restrict__ offsets,
                                                const float *__restrict__ sil_pred, float *__restrict__ vdist,
                                                float *__restrict__ d_sil, int *__restrict__ ind_ba,
                                                const float *__restrict__ gparam, const int *__restrict__ gstart,
                                                const float4 *__restrict__ sortedA, const float4 *__restrict__ sortedB,
                                                const unsigned char *__restrict__ gaux, const int *__restrict__ lat_ok) {
  __shared__ float4 s4[GRID ? 1 : MTILE];
  __shared__ __align__(16) unsigned char s_aux[GRID ? GRID_AUX_BYTES : 16];
  int i = blockIdx.y;
  if (GRID && lat_ok && lat_ok[i]) return;   // searched on the pixel lattice by k_mesh_ba_lat
  if (GRID) {   // occupancy masks + cell distances of this image's PIXEL grid
    const uint4 *src = reinterpret_cast<const uint4 *>(gaux + ((size_t)i * 2 + 1) * GRID_AUX_BYTES);
    for (int k = threadIdx.x; k < GRID_AUX_BYTES / 16; k += MT) reinterpret_cast<uint4 *>(s_aux)[k] = src[k];
    __syncthreads();
  }
  int p0 = offsets[i], np = offsets[i + 1] - p0;
  int slot = blockIdx.x * MT + threadIdx.x;
  bool ok = slot < V;
  const float *Bv = sil_pred + (size_t)i * V * 2;
  // GRID: threads walk the vertices in their binned order, so the lanes of a warp query
  // neighbouring positions (similar cells, similar trip counts); results go to the vertex's
  // original index, and the per-image sum is taken in index order by k_mesh_rowsum.
  int b = slot;
  if (GRID && ok) b = __float_as_int(sortedB[(size_t)i * V + slot].w);
  float bx = 0.f, by = 0.f;
  if (ok) {
    bx = Bv[(size_t)b * 2 + 0];
    by = Bv[(size_t)b * 2 + 1];
  }
  float b2 = __fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by));
  float best = 3.4e38f;
  int ai = -1;
  if (GRID) {
    if (ok && np > 0) {
      const float *gpB = gparam + ((size_t)i * 2 + 0) * GP_STRIDE, *gpA = gparam + ((size_t)i * 2 + 1) * GP_STRIDE;
      float margin = 32.0f * 5.9604645e-8f * fmaxf(gpA[6], gpB[6]);
      grid_search<false>(bx, by, b2, sortedA + p0, gstart + ((size_t)i * 2 + 1) * (GRID_NC + 1), gpA, margin,
                         reinterpret_cast<const unsigned long long *>(s_aux),
                         reinterpret_cast<const unsigned long long *>(s_aux) + GRID_G, s_aux + 2 * GRID_G * 8, best, ai);
    }
  } else
  for (int a0 = 0; a0 < np; a0 += MTILE) {
    int na = min(MTILE, np - a0);
    __syncthreads();
    for (int q = threadIdx.x; q < na; q += MT) {
      float ax = pts[(size_t)(p0 + a0 + q) * 2 + 0], ay = pts[(size_t)(p0 + a0 + q) * 2 + 1];
      s4[q] = make_float4(ax, ay, __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int q = 0; q < na; ++q) {
      float4 aq = s4[q];
      float d = d2_expand(aq.x, aq.y, aq.z, bx, by, b2);
      if (d < best) {
        best = d;
        ai = a0 + q;
      }
    }
  }
  float dist = 0.f, gx = 0.f, gy = 0.f;
  if (ok && ind_ba) ind_ba[(size_t)i * V + b] = ai;
  if (ok && ai >= 0) {
    float ax = pts[(size_t)(p0 + ai) * 2 + 0], ay = pts[(size_t)(p0 + ai) * 2 + 1];
    float dx = bx - ax, dy = by - ay;
    dist = sqrtf(dx * dx + dy * dy);
    if (dist > 0.f) {   // tf.norm's gradient is NaN at exactly 0 (measure zero); 0 here
      gx = dx / dist;
      gy = dy / dist;
    }
  }
  if (ok && d_sil) {
    d_sil[((size_t)i * V + b) * 2 + 0] = gx;
    d_sil[((size_t)i * V + b) * 2 + 1] = gy;
  }
  if (ok) vdist[(size_t)i * V + b] = dist;
}

// Per-image sum of the vertex->pixel distances in vertex-index order (fixed order: the same
// bits whichever order the search visited the vertices in).
__global__ void __launch_bounds__(256) k_mesh_rowsum(int V, const float *__restrict__ vdist, float *__restrict__ part) {
  __shared__ float red[256];
  int i = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < V; k += 256) s += vdist[(size_t)i * V + k];
  float tot = block_sum(s, red);
  if (threadIdx.x == 0) part[i] = tot;
}

// loss = sum(partials) / (3 + V)  (ops.py:129-130: silhouette_gt.shape[1] + silhouette_pred.shape[1]);
// d_sil = (unit vectors + integer sign sums) / (3 + V).
__global__ void __launch_bounds__(1024) k_mesh_finish(int n_ab, const float *__restrict__ part_ab, int n_ba,
                                                      const float *__restrict__ part_ba, float denom,
                                                      float *__restrict__ loss) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_ab; i += 1024) s += part_ab[i];
  for (int i = threadIdx.x; i < n_ba; i += 1024) s += part_ba[i];
  float tot = block_sum(s, red);
  if (threadIdx.x == 0) *loss = tot / denom;
}

__global__ void k_mesh_grad_finish(size_t n, float denom, const int *__restrict__ cnt, float *__restrict__ d_sil) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d_sil[i] = (d_sil[i] + (float)cnt[i]) / denom;
}

// ------------------------------------------------------------------------ gradient penalty
__device__ __forceinline__ void gp_col(int t, int &which, int &n, int &col) {
  if (t < 169) { which = 0; n = 169; col = t; }
  else if (t < 211) { which = 1; n = 42; col = t - 169; }
  else if (t < 221) { which = 2; n = 10; col = t - 211; }
  else { which = 3; n = 207; col = t - 221; }
}

// Column sums over a slab of rows: part[chunk][428].
__global__ void __launch_bounds__(448) k_gp_colsum(int M, int rows_per, const float *__restrict__ g0,
                                                   const float *__restrict__ g1, const float *__restrict__ g2,
                                                   const float *__restrict__ g3, float *__restrict__ part) {
  int t = threadIdx.x;
  if (t >= SMPLB_GP_FLOATS) return;
  int which, n, col;
  gp_col(t, which, n, col);
  const float *g = which == 0 ? g0 : which == 1 ? g1 : which == 2 ? g2 : g3;
  int m0 = blockIdx.x * rows_per, m1 = min(M, m0 + rows_per);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += g[(size_t)m * n + col];
  part[(size_t)blockIdx.x * SMPLB_GP_FLOATS + t] = s;
}

__global__ void __launch_bounds__(448) k_gp_sumparts(int nchunk, const float *__restrict__ part,
                                                     float *__restrict__ col_sums) {
  int t = threadIdx.x;
  if (t >= SMPLB_GP_FLOATS) return;
  float s = 0.f;
  for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * SMPLB_GP_FLOATS + t];
  col_sums[t] = s;
}

// norms[i] = || col_sums_i / M ||_F; penalty = sum (1 - norm_i)^2   (ops.py:155-163)
__device__ __forceinline__ void gp_norms(const float *__restrict__ col_sums, float invM, float *sq, float *norms) {
  int t = threadIdx.x;
  if (t < SMPLB_GP_FLOATS) {
    float m = col_sums[t] * invM;
    sq[t] = m * m;
  }
  __syncthreads();
  if (t < 4) {
    const int beg[5] = {0, 169, 211, 221, 428};
    float s = 0.f;
    for (int q = beg[t]; q < beg[t + 1]; ++q) s += sq[q];
    norms[t] = sqrtf(s);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(448) k_gp_final(float invM, const float *__restrict__ col_sums,
                                                  float *__restrict__ penalty) {
  __shared__ float sq[448];
  __shared__ float norms[4];
  gp_norms(col_sums, invM, sq, norms);
  if (threadIdx.x == 0) {
    float p = 0.f;
    for (int q = 0; q < 4; ++q) p += (1.0f - norms[q]) * (1.0f - norms[q]);
    *penalty = p;
  }
}

// d penalty / d g_i[m, col] = -2 (1 - n_i) mean_col / (n_i M_total), the same for every row m.
__global__ void __launch_bounds__(448) k_gp_bwd(int M, int rows_per, float invM, const float *__restrict__ col_sums,
                                                float *__restrict__ d0, float *__restrict__ d1, float *__restrict__ d2,
                                                float *__restrict__ d3) {
  __shared__ float sq[448];
  __shared__ float norms[4];
  gp_norms(col_sums, invM, sq, norms);
  int t = threadIdx.x;
  if (t >= SMPLB_GP_FLOATS) return;
  int which, n, col;
  gp_col(t, which, n, col);
  float *d = which == 0 ? d0 : which == 1 ? d1 : which == 2 ? d2 : d3;
  if (!d) return;
  float nn = norms[which];
  float coef = -2.0f * (1.0f - nn) * (col_sums[t] * invM) / nn * invM;
  int m0 = blockIdx.x * rows_per, m1 = min(M, m0 + rows_per);
  for (int m = m0; m < m1; ++m) d[(size_t)m * n + col] = coef;
}

// --------------------------------------------------------------------------------- launchers
int launch_kp_loss(smplb_ctx *c, int B, int K, const float *kp_gt, const float *kp_pred, float *dkp, float *part,
                   int *cnt) {
  LAUNCH(c, "kp_loss", cdiv(B, 128), 128, 0, k_kp_loss, B, K, kp_gt, kp_pred, dkp, part, cnt);
  return 0;
}

int launch_reduce_kp(smplb_ctx *c, int B, const float *part, const int *cnt, float *abs_sum, long long *num_present,
                     float *cnt_as_float) {
  LAUNCH(c, "reduce_kp", 1, RK_THREADS, 0, k_reduce_kp, B, part, cnt, abs_sum, num_present, cnt_as_float);
  return 0;
}

int launch_finalize_loss(smplb_ctx *c, float w_kp, float w_mesh, long long count_override, int have_mesh,
                         float *loss_parts) {
  LAUNCH(c, "finalize_loss", 1, 1, 0, k_finalize_loss, w_kp, w_mesh, count_override, have_mesh, c->ws_scal,
         c->ws_cnt64, loss_parts);
  return 0;
}

int launch_reduce_finalize(smplb_ctx *c, int B, float w_kp, float w_mesh, long long count_override, float *loss_parts) {
  (void)w_mesh;
  LAUNCH(c, "reduce_finalize_kp", 1, RK_THREADS, 0, k_reduce_finalize, B, c->ws_part, c->ws_cnt, w_kp, count_override, c->ws_scal,
         c->ws_cnt64, loss_parts);
  return 0;
}

#define MESH_AB_BLOCKS 32   // most CTAs per image of the pixel -> vertex kernels (= partial sums per image reserved)
// CTAs per image: a CTA's warps walk their 256-pixel chunks independently and meet only in the final block sum, so a few
// CTAs with several chunks each lose less to the slowest query of a chunk than one chunk per CTA (0.67 -> 0.56 ms at
// B = 1024 going from 32 to 8); small batches keep more CTAs to fill the SMs.
static inline int mesh_ab_blocks(const smplb_ctx *c, int B) {
  int n = cdiv(4 * c->num_sms, B > 0 ? B : 1);
  return n < 8 ? 8 : (n > MESH_AB_BLOCKS ? MESH_AB_BLOCKS : n);
}
int launch_mesh_loss(smplb_ctx *c, int B, int V, const float *pts, const int *offsets, int P, const float *sil_pred,
                     float *loss, float *d_sil_pred, int *cnt_scratch, float *part_scratch, int *ind_ab, int *ind_ba, bool finish_grad) {
  int n_ba_blocks = cdiv(V, MT);
  float *part_ab = part_scratch;
  const int ab_blocks = mesh_ab_blocks(c, B);
  float *part_ba = part_scratch + (size_t)B * MESH_AB_BLOCKS;
  if ((size_t)B * V > c->ws_vdist_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ws_vdist) CUDA_TRY(cudaFree(c->ws_vdist));
    c->ws_vdist = nullptr;
    CUDA_TRY(cudaMalloc((void **)&c->ws_vdist, (size_t)B * V * 4));
    c->ws_vdist_cap = (size_t)B * V;
  }
  if (d_sil_pred) CUDA_TRY(cudaMemsetAsync(cnt_scratch, 0, (size_t)B * V * 2 * sizeof(int), c->cur));
  if (c->use_mesh_grid) {
    // grid workspace: per image and set 8 floats + GRID_NC + 1 ints, sorted copies of both point sets
    const bool use_lat = c->use_mesh_lattice != 0;
    size_t need = (size_t)B * 2 * GP_STRIDE * 4 + (size_t)B * 2 * (GRID_NC + 1) * 4 + ((size_t)B * V + (size_t)P + 16) * 16 + 256 +
                  (size_t)B * 2 * GRID_AUX_BYTES + 128 + mesh_lattice_workspace(B);
    if (need > c->ws_grid_cap) {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (c->ws_grid) CUDA_TRY(cudaFree(c->ws_grid));
      c->ws_grid = nullptr;
      CUDA_TRY(cudaMalloc(&c->ws_grid, need));
      c->ws_grid_cap = need;
    }
    char *w = (char *)c->ws_grid;
    float4 *sortedB = (float4 *)w;
    float4 *sortedA = sortedB + (size_t)B * V;
    float *gparam = (float *)(sortedA + P + 16);
    int *gstart = (int *)(gparam + (size_t)B * 2 * GP_STRIDE);
    unsigned char *gaux = (unsigned char *)(((uintptr_t)(gstart + (size_t)B * 2 * (GRID_NC + 1)) + 63) & ~(uintptr_t)63);
    void *lat_ws = (void *)(((uintptr_t)(gaux + (size_t)B * 2 * GRID_AUX_BYTES) + 63) & ~(uintptr_t)63);
    // vertex -> pixel: images whose points are a row-major pixel list (the reference's where(seg > 0)) are searched on
    // the pixel lattice (k_mesh_lattice.cu); the CTAs of the binned-grid kernel return at once for those images.
    unsigned char *lat = use_lat ? (unsigned char *)lat_ws : nullptr;
    int *lat_ok = use_lat ? (int *)(lat + (size_t)B * LAT_BYTES) : nullptr;
    LAUNCH(c, "mesh_grid_build", dim3(2, B), 256, 0, k_grid_build, V, pts, offsets, sil_pred, gparam, gstart, sortedB, sortedA,
           gaux, lat, lat_ok);
    LAUNCH(c, "mesh_nn_pixel_to_vertex_grid", dim3(ab_blocks, B), MT, 0, k_mesh_ab<true>, V, pts, offsets, sil_pred,
           part_ab, d_sil_pred ? cnt_scratch : (int *)nullptr, ind_ab, gparam, gstart, sortedB, gaux);
    LAUNCH(c, "mesh_nn_vertex_to_pixel_grid", dim3(n_ba_blocks, B), MT, 0, k_mesh_ba<true>, V, pts, offsets, sil_pred,
           c->ws_vdist, d_sil_pred, ind_ba, gparam, gstart, sortedA, sortedB, gaux, (const int *)lat_ok);
    if (use_lat) TRY(launch_mesh_lattice_search(c, B, V, offsets, lat_ws, gparam, sortedB, c->ws_vdist, d_sil_pred, ind_ba));
  } else {
    LAUNCH(c, "mesh_nn_pixel_to_vertex", dim3(ab_blocks, B), MT, 0, k_mesh_ab<false>, V, pts, offsets, sil_pred,
           part_ab, d_sil_pred ? cnt_scratch : (int *)nullptr, ind_ab, (const float *)nullptr, (const int *)nullptr,
           (const float4 *)nullptr, (const unsigned char *)nullptr);
    LAUNCH(c, "mesh_nn_vertex_to_pixel", dim3(n_ba_blocks, B), MT, 0, k_mesh_ba<false>, V, pts, offsets, sil_pred,
           c->ws_vdist, d_sil_pred, ind_ba, (const float *)nullptr, (const int *)nullptr, (const float4 *)nullptr,
           (const float4 *)nullptr, (const unsigned char *)nullptr, (const int *)nullptr);
  }
  LAUNCH(c, "mesh_rowsum", B, 256, 0, k_mesh_rowsum, V, c->ws_vdist, part_ba);
  float denom = (float)(3 + V);
  LAUNCH(c, "mesh_finish", 1, 1024, 0, k_mesh_finish, B * ab_blocks, part_ab, B, part_ba, denom, loss);
  if (d_sil_pred && finish_grad) {   // (the step's k_proj_bwd applies (d_sil + cnt) / denom itself)
    size_t n = (size_t)B * V * 2;
    LAUNCH(c, "mesh_grad_finish", (unsigned)((n + 255) / 256), 256, 0, k_mesh_grad_finish, n, denom, cnt_scratch,
           d_sil_pred);
  }
  return 0;
}

#define GP_ROWS 64
int launch_gp_colsum(smplb_ctx *c, int M, const float *g0, const float *g1, const float *g2, const float *g3,
                     float *col_sums) {
  int nchunk = cdiv(M, GP_ROWS);
  LAUNCH(c, "gp_colsum", nchunk, 448, 0, k_gp_colsum, M, GP_ROWS, g0, g1, g2, g3, c->ws_gp);
  LAUNCH(c, "gp_sumparts", 1, 448, 0, k_gp_sumparts, nchunk, c->ws_gp, col_sums);
  return 0;
}

int launch_gp_final(smplb_ctx *c, long long M_total, const float *col_sums, float *penalty) {
  LAUNCH(c, "gp_final", 1, 448, 0, k_gp_final, 1.0f / (float)M_total, col_sums, penalty);
  return 0;
}

int launch_gp_bwd(smplb_ctx *c, int M, long long M_total, const float *col_sums, float *d0, float *d1, float *d2,
                  float *d3) {
  LAUNCH(c, "gp_bwd", cdiv(M, GP_ROWS), 448, 0, k_gp_bwd, M, GP_ROWS, 1.0f / (float)M_total, col_sums, d0, d1, d2, d3);
  return 0;
}
