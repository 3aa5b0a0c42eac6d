/*
 * flux_kernels.cuh -- the pseudo flux (src/flux.c:111-201), the step that consumes the exchanged gradients
 * (SURVEY 8f row f3), on the tile schedule of the gradient kernel (schedule.cpp).
 *
 * What flux.c computes, per face (p0, p1) with area vector n:
 *   d[eq][c]   = 0.5 * (grad[p0][eq][c] + grad[p1][eq][c])        eq = IVX..IVZ (0..2), c = x,y,z   (flux.c:153-163)
 *   sts_xx     = lambda * (d[1][1] + d[2][2] - 2*d[0][0]), ...     lambda = -2/3 * mue_eff, mue_eff = 1 (:165-173)
 *   flux_IVX   = -(sts_xx*nx + sts_xy*ny + sts_xz*nz), ...         (:175-177)
 *   if (ftype != 3) psd_flux[p0] += flux;  if (ftype != 2) psd_flux[p1] -= flux;               (:179-190)
 * where ftype is the face type of the colour: 1 = p0 not written by this thread, 2 = p1 not written, 3 = both
 * written (rangelist.c:719-736).  The tests differ from gradients.c (:64, :86, :108): a face between two own points
 * only updates p1, a face whose p1 is a ghost updates p0.  The GPU kernel restates exactly that for a single-thread
 * run (ftype 1 <=> p0 is a ghost, 2 <=> p1 is a ghost): an own point p receives
 *   - flux   from every face where it is p1,
 *   + flux   from the faces where it is p0 and p1 is a ghost,
 * summed from 0 in the reference's face order.  Ghost rows of psd_flux are not defined by the reference (they are
 * accumulated into without ever being zeroed) and are not written.
 *
 * One CTA per tile, two CTAs per SM.  The normals and the ELL adjacency of the tile blob arrive by two TMA bulk
 * copies; the first 72 bytes (grad[p][0..2][0..2]) of the grad rows of the tile's own and halo points are
 * gathered with 8-byte cp.async.  One thread per own point walks its ELL column; nothing is written to shared
 * memory after the loads.  ELL entry: local point | ghost << 15 | face slot << 16 | (point is p1) << 31.
 * EXACT = IEEE multiply and add in the reference's expression order (bit-identical to the reference built without
 * FMA); otherwise the compiler may contract.
 */
#ifndef CFDP_FLUX_KERNELS_CUH
#define CFDP_FLUX_KERNELS_CUH

#include "gg_kernels.cuh"

namespace ggk {

#define CFDP_FLUX_ROW 9 /* doubles of a grad row the flux reads: [IVX..IVZ][0..2] (flux.h:7-9) */

/* shared memory of a tile: [normals][ELL adjacency][grad rows]; the halo row list of the blob is not staged (it is
 * only needed to address the gather), which keeps the largest tiles below the two-CTAs-per-SM limit */
__host__ __device__ __forceinline__ uint32_t flux_adj_src(uint32_t halo_off, uint32_t nhalo) { return halo_off + ((nhalo * 4u + 15u) & ~15u); }
__host__ __device__ __forceinline__ uint32_t flux_rows_off(uint32_t blob_bytes, uint32_t halo_off, uint32_t nhalo)
{
  return halo_off + (blob_bytes - flux_adj_src(halo_off, nhalo)); /* multiples of 16 */
}
__host__ __device__ __forceinline__ uint32_t flux_footprint(uint32_t blob_bytes, uint32_t halo_off, uint32_t npts, uint32_t nhalo)
{
  return flux_rows_off(blob_bytes, halo_off, nhalo) + (CFDP_HALO_BASE(npts) + nhalo) * (CFDP_FLUX_ROW * 8);
}

template <bool EXACT>
__device__ __forceinline__ void face_flux(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ n,
                                          double &fx, double &fy, double &fz)
{
  const double lambda = -2.0 / 3.0; /* -2/3 * mue_eff, mue_eff = 1 (flux.c:123, :165) */
  if (EXACT) {
    double d[CFDP_FLUX_ROW];
#pragma unroll
    for (int k = 0; k < CFDP_FLUX_ROW; k++) d[k] = __dmul_rn(0.5, __dadd_rn(a[k], b[k]));
    /* d[0..2] = dvx_dx,dy,dz   d[3..5] = dvy_dx,dy,dz   d[6..8] = dvz_dx,dy,dz */
    const double sxx = __dmul_rn(lambda, __dadd_rn(__dadd_rn(d[4], d[8]), -__dmul_rn(2.0, d[0])));
    const double syy = __dmul_rn(lambda, __dadd_rn(__dadd_rn(d[0], d[8]), -__dmul_rn(2.0, d[4])));
    const double szz = __dmul_rn(lambda, __dadd_rn(__dadd_rn(d[0], d[4]), -__dmul_rn(2.0, d[8])));
    const double sxy = __dadd_rn(d[1], d[3]), sxz = __dadd_rn(d[2], d[6]), syz = __dadd_rn(d[5], d[7]);
    fx = -__dadd_rn(__dadd_rn(__dmul_rn(sxx, n[0]), __dmul_rn(sxy, n[1])), __dmul_rn(sxz, n[2]));
    fy = -__dadd_rn(__dadd_rn(__dmul_rn(sxy, n[0]), __dmul_rn(syy, n[1])), __dmul_rn(syz, n[2]));
    fz = -__dadd_rn(__dadd_rn(__dmul_rn(sxz, n[0]), __dmul_rn(syz, n[1])), __dmul_rn(szz, n[2]));
  } else {
    double d[CFDP_FLUX_ROW];
#pragma unroll
    for (int k = 0; k < CFDP_FLUX_ROW; k++) d[k] = 0.5 * (a[k] + b[k]);
    const double sxx = lambda * (d[4] + d[8] - 2.0 * d[0]);
    const double syy = lambda * (d[0] + d[8] - 2.0 * d[4]);
    const double szz = lambda * (d[0] + d[4] - 2.0 * d[8]);
    const double sxy = d[1] + d[3], sxz = d[2] + d[6], syz = d[5] + d[7];
    fx = -(sxx * n[0] + sxy * n[1] + sxz * n[2]);
    fy = -(sxy * n[0] + syy * n[1] + syz * n[2]);
    fz = -(sxz * n[0] + syz * n[1] + szz * n[2]);
  }
}

/* one tile per CTA (kept as a second, independent implementation: CFDP_FLUX_KERNEL=1) */
template <bool EXACT>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS, 2)
psd_flux_tile_kernel(const TileDesc *__restrict__ tiles, const unsigned char *__restrict__ blob,
                     const double *__restrict__ grad, double *__restrict__ flux)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full;
  const TileDesc td = tiles[blockIdx.x];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npts = td.npts, nhalo = td.nhalo, n_even = CFDP_HALO_BASE(npts);
  const unsigned char *tb = blob + td.blob_off();
  const uint32_t adj_src = flux_adj_src(td.halo_off, td.nhalo), adj_bytes = td.blob_bytes - adj_src;
  double *s_rows = reinterpret_cast<double *>(smem + flux_rows_off(td.blob_bytes, td.halo_off, td.nhalo));

  if (tid == 0) {
    mbar_init(&full, 1);
    fence_mbar_init();
    const uint64_t pol = l2_policy_evict_first();
    mbar_arrive_expect_tx(&full, td.halo_off + adj_bytes);
    if (td.halo_off) bulk_g2s_hint(smem, tb, td.halo_off, &full, pol);
    if (adj_bytes) bulk_g2s_hint(smem + td.halo_off, tb + adj_src, adj_bytes, &full, pol);
  }
  { /* own rows: consecutive lanes = consecutive words of a row (72 of its 168 bytes) */
    const double *g = grad + (size_t)td.row0 * (NGRAD * 3);
    const int nw = npts * CFDP_FLUX_ROW;
    for (int i = tid; i < nw; i += nthr) {
      const int r = i / CFDP_FLUX_ROW, c = i - r * CFDP_FLUX_ROW;
      cp_async8(s_rows + i, g + (size_t)r * (NGRAD * 3) + c);
    }
  }
  { /* halo rows, through the tile's halo row list (read from global memory: the blob is still in flight) */
    const uint32_t *hrows = reinterpret_cast<const uint32_t *>(tb + td.halo_off);
    double *s = s_rows + n_even * CFDP_FLUX_ROW;
    const int nw = nhalo * CFDP_FLUX_ROW;
    for (int i = tid; i < nw; i += nthr) {
      const int r = i / CFDP_FLUX_ROW, c = i - r * CFDP_FLUX_ROW;
      const uint32_t row = __ldg(hrows + r);
      if (row != 0xFFFFFFFFu) cp_async8(s + i, grad + (size_t)row * (NGRAD * 3) + c);
    }
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();          /* everybody's rows have landed; the mbarrier initialisation is visible */
  mbar_wait(&full, 0);      /* the blob has landed */

  if (tid < npts) {
    const double *s_nrm = reinterpret_cast<const double *>(smem);
    const uint32_t *ell = reinterpret_cast<const uint32_t *>(smem + td.halo_off) + tid;
    const int npad = td.npad, maxdeg = td.maxdeg;
    double a[CFDP_FLUX_ROW];
#pragma unroll
    for (int k = 0; k < CFDP_FLUX_ROW; k++) a[k] = s_rows[tid * CFDP_FLUX_ROW + k];
    double ax = 0.0, ay = 0.0, az = 0.0;
    uint32_t e_next = maxdeg > 0 ? ell[0] : CFDP_ADJ_PAD;
    for (int j = 0; j < maxdeg; j++) {
      const uint32_t e = e_next;
      e_next = j + 1 < maxdeg ? ell[(j + 1) * npad] : CFDP_ADJ_PAD;
      if (e == CFDP_ADJ_PAD) continue;
      const bool is_p1 = (e & 0x80000000u) != 0;
      if (!is_p1 && !(e & 0x8000u)) continue; /* this point is p0 and p1 is an own point: flux.c:179 skips p0 (ftype 3) */
      double fx, fy, fz;
      face_flux<EXACT>(a, s_rows + CFDP_FLUX_ROW * (e & 0x7FFFu), s_nrm + 3 * ((e >> 16) & 0x7FFFu), fx, fy, fz);
      if (is_p1) { fx = -fx; fy = -fy; fz = -fz; }   /* psd_flux[p1] -= flux (flux.c:185-190) */
      ax = __dadd_rn(ax, fx); ay = __dadd_rn(ay, fy); az = __dadd_rn(az, fz);
    }
    double *o = flux + (size_t)(td.row0 + tid) * NFLUX;
    o[0] = ax; o[1] = ay; o[2] = az;
  }
}

/* 72 bytes of a grad row -> shared memory in 16-byte pieces.  168 = 72 = 8 (mod 16): when the global row and its
 * shared-memory position have the same parity, both sides are 16-byte aligned at byte 0 (even) or at byte 8 (odd). */
__device__ __forceinline__ void flux_row_piece(double *sdst, const double *gsrc, bool odd, int k /* 0..4 */)
{
  if (!odd) {
    if (k < 4) cp_async16(sdst + 2 * k, gsrc + 2 * k); else cp_async8(sdst + 8, gsrc + 8);
  } else {
    if (k == 0) cp_async8(sdst, gsrc); else cp_async16(sdst + 2 * k - 1, gsrc + 2 * k - 1);
  }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#define CFDP_FLUX_HALO_PER_THREAD 4 /* halo rows a thread gathers: nhalo <= 4 * blockDim */
#define CFDP_FLUX_MAX_CHUNK 16       /* tiles per CTA (descriptors in static shared memory: every byte counts against two CTAs per SM) */

/*
 * Production pseudo-flux kernel: two CTAs per SM, each walking a chunk of consecutive tiles.  Per tile:
 *   load   thread 0 starts the two bulk copies (normals, adjacency); all threads gather the grad rows: own rows in
 *          16-byte pieces (consecutive lanes = consecutive pieces), halo rows one row per thread from row numbers
 *          that were fetched into registers during the previous tile's walk (no dependent global load in the way)
 *   walk   one thread per own point; the three sums go from registers straight to global memory
 * While one CTA of the SM loads, the other walks.
 */
template <bool EXACT>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS, 2)
psd_flux_pipe_kernel(const TileDesc *__restrict__ tiles, int ntiles, int chunk, const unsigned char *__restrict__ blob,
                     const double *__restrict__ grad, double *__restrict__ flux, int variant)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full;
  __shared__ TileDesc s_tds[CFDP_FLUX_MAX_CHUNK];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int t_begin = blockIdx.x * chunk;
  const int t_end = min(t_begin + chunk, ntiles);
  if (t_begin >= t_end) return;
  {
    const int nw = (t_end - t_begin) * (int)(sizeof(TileDesc) / 4);
    const uint32_t *g = reinterpret_cast<const uint32_t *>(tiles + t_begin);
    uint32_t *d = reinterpret_cast<uint32_t *>(s_tds);
    for (int i = tid; i < nw; i += nthr) d[i] = __ldg(g + i);
  }
  if (tid == 0) {
    mbar_init(&full, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint64_t pol = l2_policy_evict_first();

  uint32_t hrow[CFDP_FLUX_HALO_PER_THREAD]; /* device rows of this thread's halo positions tid, tid + nthr, ... of the next tile to load */
  auto fetch_halo_rows = [&](int t) {
    if (t >= t_end) return;
    const TileDesc pd = s_tds[t - t_begin];
    const uint32_t *g = reinterpret_cast<const uint32_t *>(blob + pd.blob_off() + pd.halo_off);
#pragma unroll
    for (int k = 0; k < CFDP_FLUX_HALO_PER_THREAD; k++) {
      const int h = tid + k * nthr;
      hrow[k] = h < (int)pd.nhalo ? __ldg(g + h) : 0xFFFFFFFFu;
    }
  };
  fetch_halo_rows(t_begin);

  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const TileDesc td = s_tds[t - t_begin];
    const int npts = td.npts, n_even = CFDP_HALO_BASE(npts);
    const unsigned char *tb = blob + td.blob_off();
    const uint32_t adj_src = flux_adj_src(td.halo_off, td.nhalo), adj_bytes = td.blob_bytes - adj_src;
    double *s_rows = reinterpret_cast<double *>(smem + flux_rows_off(td.blob_bytes, td.halo_off, td.nhalo));
    if (tid == 0) {
      mbar_arrive_expect_tx(&full, td.halo_off + adj_bytes);
      if (td.halo_off) bulk_g2s_hint(smem, tb, td.halo_off, &full, pol);
      if (adj_bytes) bulk_g2s_hint(smem + td.halo_off, tb + adj_src, adj_bytes, &full, pol);
    }
    { /* own rows: row0 is even and so is the first shared-memory row */
      const double *g = grad + (size_t)td.row0 * (NGRAD * 3);
      const int np = npts * 5;
      for (int i = tid; i < np; i += nthr) {
        const int r = i / 5, k = i - r * 5;
        flux_row_piece(s_rows + r * CFDP_FLUX_ROW, g + (size_t)r * (NGRAD * 3), r & 1, k);
      }
    }
    { /* halo rows: one row per thread */
      double *s = s_rows + n_even * CFDP_FLUX_ROW;
#pragma unroll
      for (int k = 0; k < CFDP_FLUX_HALO_PER_THREAD; k++) {
        const uint32_t row = hrow[k];
        if (row == 0xFFFFFFFFu) continue;
        const int h = tid + k * nthr;
        double *sd = s + h * CFDP_FLUX_ROW;
        const double *gs = grad + (size_t)row * (NGRAD * 3);
        if (((row ^ (uint32_t)h) & 1u) == 0) {
#pragma unroll
          for (int q = 0; q < 5; q++) flux_row_piece(sd, gs, row & 1u, q);
        } else {
#pragma unroll
          for (int q = 0; q < CFDP_FLUX_ROW; q++) cp_async8(sd + q, gs + q);
        }
      }
    }
    cp_async_commit();
    fetch_halo_rows(t + 1);   /* consumed by the next iteration: the latency hides behind this tile's wait and walk */
    cp_async_wait_all();
    __syncthreads();                       /* everybody's rows have landed */
    mbar_wait(&full, (uint32_t)it & 1u);   /* normals and adjacency have landed */

    if (variant && t + 1 < t_end) {
      /* everything the next tile reads is asked into L2 before this tile's walk: its loads, issued after the walk and
       * waited for at once, then cost an L2 round trip instead of a DRAM one.  1 = blob, 2 = grad rows, 3 = both */
      const TileDesc pd = s_tds[t + 1 - t_begin];
      if ((variant & 1) && tid == 0) {
        const unsigned char *nb = blob + pd.blob_off();
        const uint32_t a_src = flux_adj_src(pd.halo_off, pd.nhalo);
        if (pd.halo_off) bulk_prefetch_l2(nb, (pd.halo_off + 15u) & ~15u);
        if (pd.blob_bytes > a_src) bulk_prefetch_l2(nb + a_src, (pd.blob_bytes - a_src + 15u) & ~15u);
      }
      if (variant & 2) {
        if (tid < (int)pd.npts) {
          const char *g = reinterpret_cast<const char *>(grad + (size_t)(pd.row0 + tid) * (NGRAD * 3));
          prefetch_l2(g); prefetch_l2(g + 64);
        }
#pragma unroll
        for (int k = 0; k < CFDP_FLUX_HALO_PER_THREAD; k++)
          if (hrow[k] != 0xFFFFFFFFu) {
            const char *g = reinterpret_cast<const char *>(grad + (size_t)hrow[k] * (NGRAD * 3));
            prefetch_l2(g); prefetch_l2(g + 64);
          }
      }
    }

    if (tid < npts) {
      const double *s_nrm = reinterpret_cast<const double *>(smem);
      const uint32_t *ell = reinterpret_cast<const uint32_t *>(smem + td.halo_off) + tid;
      const int npad = td.npad, maxdeg = td.maxdeg;
      double a[CFDP_FLUX_ROW];
#pragma unroll
      for (int k = 0; k < CFDP_FLUX_ROW; k++) a[k] = s_rows[tid * CFDP_FLUX_ROW + k];
      double ax = 0.0, ay = 0.0, az = 0.0;
      uint32_t e_next = maxdeg > 0 ? ell[0] : CFDP_ADJ_PAD;
      for (int j = 0; j < maxdeg; j++) {
        const uint32_t e = e_next;
        e_next = j + 1 < maxdeg ? ell[(j + 1) * npad] : CFDP_ADJ_PAD;
        if (e == CFDP_ADJ_PAD) continue;
        const bool is_p1 = (e & 0x80000000u) != 0;
        if (!is_p1 && !(e & 0x8000u)) continue; /* this point is p0 and p1 is an own point: flux.c:179 skips p0 (ftype 3) */
        double fx, fy, fz;
        face_flux<EXACT>(a, s_rows + CFDP_FLUX_ROW * (e & 0x7FFFu), s_nrm + 3 * ((e >> 16) & 0x7FFFu), fx, fy, fz);
        if (is_p1) { fx = -fx; fy = -fy; fz = -fz; }   /* psd_flux[p1] -= flux (flux.c:185-190) */
        ax = __dadd_rn(ax, fx); ay = __dadd_rn(ay, fy); az = __dadd_rn(az, fz);
      }
      double *o = flux + (size_t)(td.row0 + tid) * NFLUX;
      o[0] = ax; o[1] = ay; o[2] = az;
    }
    __syncthreads(); /* shared memory is free for the next tile */
  }
}

} // namespace ggk
#endif
