/*
 * gg_kernels.cuh -- the Green-Gauss tile kernels (sm_100a).
 *
 * Replaces private_compute_gradients_gg (reference src/gradients.c:25-147): zero at first touch,
 * per face  val = 0.5*(var[p0]+var[p1]);  grad[p0] += n*val;  grad[p1] -= n*val,  scale by
 * 1/pvolume at last touch.  Here every own point belongs to exactly one tile; one thread owns
 * the point, walks its incident faces in the reference's single-thread order and keeps the 7x3
 * sums in registers: no atomics, no read of grad, one 168-byte row store per point.
 *
 * gg_tile_pipe_kernel (the production kernel): a CTA processes a chunk of consecutive tiles
 * through a two-stage shared-memory ring.  Per tile, ONE elected thread issues three TMA bulk
 * copies (cp.async.bulk global->shared, mbarrier complete_tx): the tile blob (face normals read
 * once, halo row list, ELL adjacency), the contiguous var rows and volumes of the tile's own
 * points; all threads gather the var rows of the tile's halo points with 8-byte cp.async
 * (LDGSTS) tracked by the same mbarrier.  The loads of tile t+2 are issued when tile t retires,
 * so they fly while tile t+1 computes; HBM is only touched by asynchronous copies and by the
 * coalesced 16-byte row stores.
 *
 * Arithmetic modes
 *   EXACT = true : separate IEEE multiply and add in the reference's order.  The device holds
 *                  hvar = 0.5*var (halved once when var is uploaded): 0.5*(a+b) == 0.5*a + 0.5*b bit
 *                  for bit (power-of-two scaling commutes with rounding; subnormal inputs
 *                  excepted) -> bit-identical to the reference built without FMA, one thread.
 *   EXACT = false: fused multiply-add (more accurate, not bit-identical).
 */
#ifndef CFDP_GG_KERNELS_CUH
#define CFDP_GG_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "common.h"

#define CFDP_HALO_PER_THREAD 8 /* halo rows a thread can gather per tile: nhalo <= 8 * blockDim */
#define CFDP_MAX_CHUNK 64      /* tiles per CTA */

namespace ggk {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double flip_sign(double x, uint32_t signbit)
{
  return __hiloint2double(__double2hiint(x) ^ (int)signbit, __double2loint(x));
}

/* the face walk of one point for the equations [LO, LO+CNT): ell = this thread's ELL column */
template <bool EXACT, int LO, int CNT>
__device__ __forceinline__ void walk_faces(const uint32_t *__restrict__ ell, int npad, int maxdeg, const double *__restrict__ s_nrm,
                                           const double *__restrict__ s_hvar, const double (&hv)[CNT], double (&acc)[CNT * 3])
{
#pragma unroll 2
  for (int j = 0; j < maxdeg; j++) {
    const uint32_t e = ell[j * npad];
    if (e == CFDP_ADJ_PAD) continue;
    const double *n = s_nrm + 3 * ((e >> 16) & 0x7FFFu);
    const double *w = s_hvar + NGRAD * (e & 0xFFFFu) + LO;
    const uint32_t sb = e & 0x80000000u;      /* this point is p1 of the face: grad[p1] -= n*val (gradients.c:101-105) */
    const double nx = flip_sign(n[0], sb), ny = flip_sign(n[1], sb), nz = flip_sign(n[2], sb);
#pragma unroll
    for (int q = 0; q < CNT; q++) {
      if (EXACT) {
        const double val = __dadd_rn(hv[q], w[q]);                     /* == 0.5*(var[p0]+var[p1]), gradients.c:77 */
        acc[3 * q + 0] = __dadd_rn(acc[3 * q + 0], __dmul_rn(nx, val));
        acc[3 * q + 1] = __dadd_rn(acc[3 * q + 1], __dmul_rn(ny, val));
        acc[3 * q + 2] = __dadd_rn(acc[3 * q + 2], __dmul_rn(nz, val));
      } else {
        const double val = hv[q] + w[q];
        acc[3 * q + 0] = fma(nx, val, acc[3 * q + 0]);
        acc[3 * q + 1] = fma(ny, val, acc[3 * q + 1]);
        acc[3 * q + 2] = fma(nz, val, acc[3 * q + 2]);
      }
    }
  }
}

/* one point, equations [LO, LO+CNT): walk, scale by 1/volume, park the partial row in the output staging */
template <bool EXACT, int LO, int CNT>
__device__ __forceinline__ void point_rows(int p, const uint32_t *__restrict__ ell0, int npad, int maxdeg, const double *__restrict__ s_nrm,
                                           const double *__restrict__ s_hvar, double inv_vol, double (&acc)[CNT * 3])
{
  double hv[CNT];
#pragma unroll
  for (int q = 0; q < CNT; q++) hv[q] = s_hvar[p * NGRAD + LO + q];
#pragma unroll
  for (int k = 0; k < CNT * 3; k++) acc[k] = 0.0;
  walk_faces<EXACT, LO, CNT>(ell0 + p, npad, maxdeg, s_nrm, s_hvar, hv, acc);
#pragma unroll
  for (int k = 0; k < CNT * 3; k++) acc[k] = __dmul_rn(acc[k], inv_vol);
}

/* shared-memory stage of a tile: [blob | output rows (aliased)][half-var rows][volumes], each part 128-byte aligned;
 * the offsets depend on the tile, stage_bytes is the largest footprint of any tile */
__host__ __device__ __forceinline__ uint32_t tile_var_off(uint32_t blob_bytes, uint32_t npts)
{
  const uint32_t out_bytes = npts * (NGRAD * 3 * 8);
  return ((blob_bytes > out_bytes ? blob_bytes : out_bytes) + 127u) & ~127u;
}
__host__ __device__ __forceinline__ uint32_t tile_pvol_off(uint32_t blob_bytes, uint32_t npts, uint32_t nhalo)
{
  return tile_var_off(blob_bytes, npts) + (((CFDP_HALO_BASE(npts) + nhalo) * (NGRAD * 8) + 127u) & ~127u);
}
__host__ __device__ __forceinline__ uint32_t tile_footprint(uint32_t blob_bytes, uint32_t npts, uint32_t nhalo)
{
  return tile_pvol_off(blob_bytes, npts, nhalo) + ((CFDP_HALO_BASE(npts) * 8 + 127u) & ~127u);
}
struct PipeLayout {
  uint32_t stage_bytes;
  int stages;        /* 1: one buffer, latency hidden by a second resident CTA; 2: double buffered inside the CTA */
  int block_points;  /* threads per equation group (multiple of 32, >= largest tile) */
};

/*
 * SPLIT = 1: one thread per point, 21 sums in registers.
 * SPLIT = 2: two threads per point (equations 0-3 and 4-6, warp-uniform roles): twice the warps for
 *            latency hiding at half the registers; normals / adjacency are read by both.
 * MINB     : resident CTAs per SM the register allocation must allow (2 with single-stage staging).
 */
template <bool EXACT, int SPLIT, int MINB>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS * SPLIT, MINB)
gg_tile_pipe_kernel(const TileDesc *__restrict__ tiles, int ntiles, int chunk, const unsigned char *__restrict__ blob,
                    const double *__restrict__ hvar, const double *__restrict__ pvol, double *__restrict__ grad, PipeLayout L)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[2];
  __shared__ TileDesc s_tds[CFDP_MAX_CHUNK];   /* descriptors of this CTA's tiles */
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int t_begin = blockIdx.x * chunk;
  const int t_end = min(t_begin + chunk, ntiles);
  if (t_begin >= t_end) return;
  const int nst = L.stages;
  {
    const int nw = (t_end - t_begin) * (int)(sizeof(TileDesc) / 4);
    const uint32_t *g = reinterpret_cast<const uint32_t *>(tiles + t_begin);
    uint32_t *d = reinterpret_cast<uint32_t *>(s_tds);
    for (int i = tid; i < nw; i += nthr) d[i] = __ldg(g + i);
  }

  if (tid == 0) {
    mbar_init(&full[0], (uint32_t)nthr + 1);
    mbar_init(&full[1], (uint32_t)nthr + 1);
    fence_mbar_init();
  }

  /* prefetch state of this thread: descriptor and halo rows of the tile to be fetched next */
  TileDesc pf_td;
  uint32_t pf_h[CFDP_HALO_PER_THREAD];
  auto load_pf_meta = [&](int t) {
    if (t < t_end) {
      pf_td = s_tds[t - t_begin];
      const uint32_t *hr = reinterpret_cast<const uint32_t *>(blob + pf_td.blob + pf_td.halo_off);
#pragma unroll
      for (int k = 0; k < CFDP_HALO_PER_THREAD; k++) {
        const int i = tid + k * nthr;
        pf_h[k] = i < (int)pf_td.nhalo ? __ldg(hr + i) : 0xFFFFFFFFu;
      }
    }
  };
  auto issue_pf = [&](int t, int s) {
    if (t < t_end) {
      unsigned char *st = smem + (size_t)s * L.stage_bytes;
      const uint32_t n_even = CFDP_HALO_BASE((uint32_t)pf_td.npts);
      const uint32_t voff = tile_var_off(pf_td.blob_bytes, pf_td.npts);
      if (tid == 0) {
        const uint32_t nb = pf_td.blob_bytes, nv = n_even * (NGRAD * 8), np = n_even * 8;
        mbar_arrive_expect_tx(&full[s], nb + nv + np);
        bulk_g2s(st, blob + pf_td.blob, nb, &full[s]);
        bulk_g2s(st + voff, hvar + (size_t)pf_td.row0 * NGRAD, nv, &full[s]);
        bulk_g2s(st + tile_pvol_off(pf_td.blob_bytes, pf_td.npts, pf_td.nhalo), pvol + pf_td.row0, np, &full[s]);
      }
      double *vs = reinterpret_cast<double *>(st + voff) + (size_t)n_even * NGRAD;
#pragma unroll
      for (int k = 0; k < CFDP_HALO_PER_THREAD; k++) {
        if (pf_h[k] != 0xFFFFFFFFu) {
          const double *src = hvar + (size_t)pf_h[k] * NGRAD;
          double *dst = vs + (size_t)(tid + k * nthr) * NGRAD;
#pragma unroll
          for (int c = 0; c < NGRAD; c++) cp_async8(dst + c, src + c);
        }
      }
      cp_async_mbar_arrive_noinc(&full[s]);
    }
  };

  __syncthreads(); /* descriptors and mbarriers visible */
  load_pf_meta(t_begin);
  issue_pf(t_begin, 0);
  if (nst == 2) { load_pf_meta(t_begin + 1); issue_pf(t_begin + 1, 1); }

  const int grp = tid / L.block_points;         /* warp-uniform: block_points is a multiple of 32 */
  const int p = tid - grp * L.block_points;

  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const int s = nst == 2 ? (it & 1) : 0;
    const uint32_t parity = (uint32_t)(nst == 2 ? (it >> 1) : it) & 1u;
    load_pf_meta(t + nst); /* consumed when this tile retires: latency hidden behind the face walk */
    unsigned char *st = smem + (size_t)s * L.stage_bytes;
    const TileDesc td = s_tds[t - t_begin];
    const int npts = td.npts, nhalo = td.nhalo;
    double *s_nrm = reinterpret_cast<double *>(st);
    const double *s_hvar = reinterpret_cast<const double *>(st + tile_var_off(td.blob_bytes, td.npts));
    const double *s_pvol = reinterpret_cast<const double *>(st + tile_pvol_off(td.blob_bytes, td.npts, td.nhalo));
    mbar_wait(&full[s], parity);
    const uint32_t *ell0 = reinterpret_cast<const uint32_t *>(st + td.halo_off + ((nhalo * 4 + 15) & ~15));

    double acc[(SPLIT == 1 ? NGRAD : 4) * 3];
    const bool active = p < npts;
    if (active) {
      const double inv_vol = __ddiv_rn(1.0, s_pvol[p]);                 /* gradients.c:138 */
      if (SPLIT == 1) {
        point_rows<EXACT, 0, NGRAD>(p, ell0, td.npad, td.maxdeg, s_nrm, s_hvar, inv_vol, reinterpret_cast<double(&)[NGRAD * 3]>(acc));
      } else if (grp == 0) {
        point_rows<EXACT, 0, 4>(p, ell0, td.npad, td.maxdeg, s_nrm, s_hvar, inv_vol, reinterpret_cast<double(&)[12]>(acc));
      } else {
        point_rows<EXACT, 4, 3>(p, ell0, td.npad, td.maxdeg, s_nrm, s_hvar, inv_vol, reinterpret_cast<double(&)[9]>(acc));
      }
    }
    __syncthreads(); /* normals and adjacency are dead: the blob region becomes the output staging */
    if (active) {
      double *o = s_nrm + p * (NGRAD * 3);
      if (SPLIT == 1) {
#pragma unroll
        for (int k = 0; k < NGRAD * 3; k++) o[k] = acc[k];
      } else if (grp == 0) {
#pragma unroll
        for (int k = 0; k < 12; k++) o[k] = acc[k];
      } else {
#pragma unroll
        for (int k = 0; k < 9; k++) o[12 + k] = acc[k];
      }
    }
    __syncthreads();
    {
      double *gout = grad + (size_t)td.row0 * (NGRAD * 3);
      const int n = npts * NGRAD * 3, n2 = n >> 1;
      double2 *g2 = reinterpret_cast<double2 *>(gout);
      const double2 *s2 = reinterpret_cast<const double2 *>(s_nrm);
      for (int i = tid; i < n2; i += nthr) g2[i] = s2[i];
      if ((n & 1) && tid == 0) gout[n - 1] = s_nrm[n - 1];
    }
    fence_proxy_async(); /* generic-proxy accesses of this stage are ordered before the next bulk copy into it */
    __syncthreads();
    issue_pf(t + nst, s);
  }
}

/* v1: one tile per CTA, synchronous staging (kept as a second, independent implementation) */
template <bool EXACT>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS, 2)
gg_tile_kernel(const TileDesc *__restrict__ tiles, const unsigned char *__restrict__ blob,
               const double *__restrict__ var /* 0.5 * var */, const double *__restrict__ pvol, double *__restrict__ grad,
               int region0_doubles)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_r0 = reinterpret_cast<double *>(smem_raw);  /* normals, later the output rows */
  double *s_var = s_r0 + region0_doubles;               /* [npts(even) + nhalo][7] */
  const TileDesc td = tiles[blockIdx.x];
  const unsigned char *tb = blob + td.blob;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npts = td.npts, nhalo = td.nhalo, nfaces = td.nfaces;
  const int n_even = CFDP_HALO_BASE(npts);

  { /* normals */
    const double2 *g = reinterpret_cast<const double2 *>(tb);
    double2 *s = reinterpret_cast<double2 *>(s_r0);
    const int n2 = (nfaces * 3 + 1) >> 1;
    for (int i = tid; i < n2; i += nthr) s[i] = __ldg(g + i);
  }
  { /* var rows of the tile's own points: contiguous, 16-byte aligned (row0 % 16 == 0) */
    const double2 *g = reinterpret_cast<const double2 *>(var + (size_t)td.row0 * NGRAD);
    double2 *s = reinterpret_cast<double2 *>(s_var);
    const int n2 = (n_even * NGRAD) >> 1;
    for (int i = tid; i < n2; i += nthr) s[i] = __ldg(g + i);
  }
  { /* var rows of the tile's halo points */
    const uint32_t *hrows = reinterpret_cast<const uint32_t *>(tb + td.halo_off);
    double *s = s_var + n_even * NGRAD;
    const int n = nhalo * NGRAD;
    for (int i = tid; i < n; i += nthr) {
      const int r = i / NGRAD, c = i - r * NGRAD;
      const uint32_t row = __ldg(hrows + r);
      if (row != 0xFFFFFFFFu) s[i] = __ldg(var + (size_t)row * NGRAD + c);
    }
  }
  __syncthreads();

  double acc[NGRAD * 3];
#pragma unroll
  for (int k = 0; k < NGRAD * 3; k++) acc[k] = 0.0;
  if (tid < npts) {
    double hv[NGRAD];
#pragma unroll
    for (int q = 0; q < NGRAD; q++) hv[q] = s_var[tid * NGRAD + q];
    const uint32_t *ell = reinterpret_cast<const uint32_t *>(tb + td.halo_off + ((nhalo * 4 + 15) & ~15)) + tid;
    walk_faces<EXACT, 0, NGRAD>(ell, td.npad, td.maxdeg, s_r0, s_var, hv, acc);   /* adjacency straight from global memory */
    const double tmp = __ddiv_rn(1.0, __ldg(pvol + td.row0 + tid));
#pragma unroll
    for (int k = 0; k < NGRAD * 3; k++) acc[k] = __dmul_rn(acc[k], tmp);
  }
  __syncthreads();
  if (tid < npts) {
#pragma unroll
    for (int k = 0; k < NGRAD * 3; k++) s_r0[tid * (NGRAD * 3) + k] = acc[k];
  }
  __syncthreads();
  {
    double *gout = grad + (size_t)td.row0 * (NGRAD * 3);
    const int n = npts * NGRAD * 3, n2 = n >> 1;
    double2 *g2 = reinterpret_cast<double2 *>(gout);
    const double2 *s2 = reinterpret_cast<const double2 *>(s_r0);
    for (int i = tid; i < n2; i += nthr) g2[i] = s2[i];
    if ((n & 1) && tid == 0) gout[n - 1] = s_r0[n - 1];
  }
}

} // namespace ggk
#endif
